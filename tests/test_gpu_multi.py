"""Several GPUs (skipped on a one-GPU box): one process driving two devices (`--devices`, mm_index_replicate over
ncclCommInitAll, parts dealt round robin to per-device pipelines), and two processes with one GPU each (the product's own
NCCL communicator: mm_comm_create, mm_index_broadcast, mm_records_allgather). The output must not depend on either."""
import os
import subprocess
import sys

import numpy as np
import pytest

import datasets
from conftest import have_gpu
from mashmap_b200 import hostlib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _n_gpus():
    try:
        import torch

        return torch.cuda.device_count()
    except Exception:  # noqa: BLE001
        return 0


pytestmark = [pytest.mark.gpu, pytest.mark.skipif(not have_gpu() or _n_gpus() < 2, reason="needs two GPUs")]


def _run(cmd, **kw):
    p = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, **kw)
    assert p.returncode == 0, (cmd, p.stderr[-3000:])
    return p


def test_devices_option_does_not_change_output(workdir):
    """mashmap-b200 --devices 0,1 == mashmap-b200 on one GPU, map and one-to-one filter modes"""
    d = datasets.make_panel_set(workdir, tag="mgp")
    for extra in ([], ["-f", "one-to-one"]):
        outs = []
        for dev in (["--device", "0"], ["--devices", "0,1"], ["--devices", "0-1", "--subBatchBases", "30000"]):
            o = os.path.join(workdir, f"mg_{len(outs)}_{len(extra)}.paf")
            _run([hostlib.CLI_PATH, "-r", d["ref"], "-q", d["qry"], "-s", "5000", "--pi", "85", "-t", "8", "-o", o] + dev + extra)
            outs.append(open(o).read())
        assert len(outs[0]) > 0
        assert outs[0] == outs[1] == outs[2]


def test_two_ranks_broadcast_and_gather(workdir):
    """bench.py under torchrun with 2 ranks on a reduced one-to-one workload: rank 0 builds the index, mm_index_broadcast hands
    it to rank 1, each rank maps its block of the reads, mm_records_allgather + the run-wide one-to-one step on rank 0;
    bench.py itself compares the sharded PAF with the single-GPU PAF of the same reads (sharded_check)."""
    import json

    env = dict(os.environ, MASTER_ADDR="127.0.0.1")
    p = _run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
              "--master-port", "29533", os.path.join(ROOT, "bench.py"), "--gpus", "2", "--config", "4", "--reads", "4000", "--ref-bp", "40000000",
              "--contigs", "8", "--steps", "1", "--warmup", "1", "--no-cpu-baseline"], env=env, cwd=ROOT)
    line = [l for l in p.stdout.splitlines() if l.startswith("{")][-1]
    j = json.loads(line)
    assert j["n_gpus"] == 2
    chk = j.get("sharded_check")
    assert chk and chk["paf_equal_to_single_gpu"], chk
