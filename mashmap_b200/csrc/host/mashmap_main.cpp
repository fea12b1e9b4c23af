/*
 * mashmap_main.cpp -- the driver program: parse -> skch::Sketch -> skch::Map, with the two timers the
 * reference prints (reference src/map/mash_map.cpp:23-57). The class API and the PAF output are the
 * drop-in boundary; the mapping itself runs on the GPU behind include/mashmap_b200.h.
 */
#include <chrono>
#include <cstdlib>
#include <iostream>

#include "skch_args.hpp"
#include "skch_index.hpp"
#include "skch_map.hpp"

int main(int argc, char **argv)
{
  unsetenv((char *)"MALLOC_ARENA_MAX");
  skch::Parameters parameters;
  skch::parseandSave(argc, argv, parameters);

  auto t0 = std::chrono::steady_clock::now();
  skch::Sketch referSketch(parameters);
  std::chrono::duration<double> timeRefSketch = std::chrono::steady_clock::now() - t0;
  std::cerr << "[mashmap-b200::map] time spent computing the reference index: " << timeRefSketch.count() << " sec" << std::endl;

  t0 = std::chrono::steady_clock::now();
  skch::Map mapper(parameters, referSketch);
  std::chrono::duration<double> timeMapQuery = std::chrono::steady_clock::now() - t0;
  std::cerr << "[mashmap-b200::map] time spent mapping the query: " << timeMapQuery.count() << " sec"
            << " (device calls " << mapper.secondsDevice << " s, host tail " << mapper.secondsHostTail << " s, input "
            << mapper.secondsInput << " s; " << mapper.totalQueryBases / timeMapQuery.count() / 1e9 << " Gbp/s)" << std::endl;
  std::cerr << "[mashmap-b200::map] mapping results saved in: " << parameters.outFileName << std::endl;
  return 0;
}
