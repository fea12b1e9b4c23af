"""Run the device pipeline many times on one batch and compare every run with the first one, per segment
(sketch, candidates, loci), for both kernel paths. Usage: python scripts/stress_determinism.py [iters]"""
import os
import sys
import tempfile

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", "tests"))
sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
import datasets  # noqa: E402
import refh  # noqa: E402
from test_gpu_stages import build_segments, upload_reference_index  # noqa: E402
from mashmap_b200 import capi  # noqa: E402


def canon(seg_res, cands, loci):
    out = []
    for i in range(len(seg_res)):
        sr = seg_res[i]
        c = cands[sr["first_candidate"]: sr["first_candidate"] + sr["n_candidates"]]
        item = [int(sr["sketch_size"]), int(sr["n_points"]), int(sr["minimum_hits"]), int(sr["best_intersection"])]
        for cc in c:
            item.append(tuple(int(cc[f]) for f in ("seqId", "rangeStartPos", "rangeEndPos", "intersectionSize")))
            l = loci[cc["first_locus"]: cc["first_locus"] + cc["n_loci"]]
            item.append(tuple(map(tuple, l.tolist())))
        out.append(item)
    return out


def main():
    iters = int(sys.argv[1]) if len(sys.argv) > 1 else 40
    wd = tempfile.mkdtemp()
    d = datasets.make_random_set(wd)
    for args in (["-s", "5000", "--pi", "95", "--dense"], ["-s", "5000", "--pi", "85"]):
        R = refh.RefSession(["-r", d["ref"], "-q", d["qry"], "-t", "4"] + args)
        for mode in ("fast", "general"):
            for v in ("MM_L1_CTA", "MM_L2_GENERAL"):
                os.environ.pop(v, None)
                if mode == "general":
                    os.environ[v] = "1"
            ctx = capi.Context(kmer_size=R.p.kmerSize, seg_length=R.p.segLength, sketch_size=R.p.sketchSize,
                               stage1_topani_filter=bool(R.p.stage1_topANI_filter))
            upload_reference_index(ctx, R)
            bases, segs, ridx, start, length = build_segments(d, R.p.segLength, R.p.kmerSize)
            ctx.batch_upload(bases, segs)
            first = None
            nbad = 0
            for it in range(iters):
                ctx.map_resident()
                cur = canon(*ctx.batch_fetch())
                sk, cnt = ctx.batch_fetch_sketch()
                cur_sk = [sk[i][: cnt[i]].tobytes() for i in range(len(cnt))]
                if first is None:
                    first, first_sk = cur, cur_sk
                    continue
                for i in range(len(cur)):
                    if cur_sk[i] != first_sk[i]:
                        nbad += 1
                        print(f"{args} {mode} iter {it} seg {i}: SKETCH differs (n {cnt[i]})")
                    elif cur[i] != first[i]:
                        nbad += 1
                        print(f"{args} {mode} iter {it} seg {i}: differs\n   first {str(first[i])[:600]}\n   now   {str(cur[i])[:600]}")
            print(f"{args} S={R.p.sketchSize} {mode}: {iters} runs x {len(first)} segments, {nbad} differences", flush=True)
            ctx.close()
        R.close()


if __name__ == "__main__":
    main()
