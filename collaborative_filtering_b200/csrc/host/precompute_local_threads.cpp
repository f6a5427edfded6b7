// precompute_local_threads n_threads -- drop-in for precompute_local_threads.cpp:215-317.  The
// reference's pool size is the number of host threads that format records here; the per-user tasks
// themselves (compute_eigens, :100-213) are batched on the GPU.
#include "precompute_common.hpp"
int main(int argc, const char** argv) {
    if (argc < 2) {
        printf("Usage:\n%s n_threads\n", argv[0]);      // :217-220
        return 1;
    }
    return gsihost::precompute_main(std::max(1, atoi(argv[1])));
}
