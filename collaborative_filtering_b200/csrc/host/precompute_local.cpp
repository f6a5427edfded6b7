// precompute_local -- drop-in for the reference tool of the same name (precompute_local.cpp:84):
// no arguments (extra ones are ignored, cf. run_test_precompute.sh:17), reads movielens/*.validate
// and ./out_fin_*, writes out_eigen_.  Serial host formatting; the math runs on the GPU.
#include "precompute_common.hpp"
int main(int, const char**) { return gsihost::precompute_main(1); }
