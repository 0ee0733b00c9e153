// bmbs-index: command-line front end of build_index.hpp (CPU data-prep tool).
#include <cstdio>
#include <cstdlib>
#include "build_index.hpp"

int main(int argc, char** argv) {
  if (argc < 2) { fprintf(stderr, "usage: bmbs-index genome.fa [threads]\n"); return 2; }
  return bmbs::build_index(argv[1], argc > 2 ? atoi(argv[2]) : 0);
}
