/*
 * skch_args.hpp -- command line -> skch::Parameters, same option names, defaults and derived values as the
 * reference's parseandSave (reference src/map/include/parseCmdArgs.hpp:257-659): auto sketch size (:620-641),
 * block_length / chain_gap default to the segment length (:474,:488), --skipSelf handling (:326-345), ...
 * Only the parsing itself is new (the reference uses a third-party ArgvParser).
 */
#ifndef SKCH_ARGS_HPP
#define SKCH_ARGS_HPP

#include "skch_types.hpp"

namespace skch {
/* exits with status 1 on a usage error, 0 after --version / --help, like the reference */
void parseandSave(int argc, char **argv, Parameters &parameters);
void printCmdOptions(const Parameters &parameters);
}  // namespace skch
#endif
