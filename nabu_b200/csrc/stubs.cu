// Entry points declared in include/nabu_b200.h whose kernels are not built yet: they fail loudly.
#include "common.cuh"
#include "nabu_b200.h"
using namespace nabu;
#define NOT_BUILT(name) do { set_error(name ": not built in this revision"); return 3; } while (0)

extern "C" size_t nabu_ctc_beam_workspace_bytes(int, int, int, int) { return 0; }
extern "C" int nabu_ctc_beam_search(const float*, const int*, int, int, int, int, int, int*, int*, float*, void*,
                                    size_t, void*) {
  NOT_BUILT("nabu_ctc_beam_search");
}
