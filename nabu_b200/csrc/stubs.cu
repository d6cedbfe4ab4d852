// Entry points declared in include/nabu_b200.h whose kernels are not built yet: they fail loudly.
#include "common.cuh"
#include "nabu_b200.h"
using namespace nabu;
#define NOT_BUILT(name) do { set_error(name ": not built in this revision"); return 3; } while (0)

extern "C" size_t nabu_speller_workspace_bytes(const nabu_speller_desc_t*) { return 0; }
extern "C" size_t nabu_speller_saved_bytes(const nabu_speller_desc_t*) { return 0; }
extern "C" int nabu_speller_fwd(const nabu_speller_desc_t*, const nabu_speller_params_t*, const float*, const int*,
                                const int*, int, const int*, float*, void*, void*, size_t, void*) {
  NOT_BUILT("nabu_speller_fwd");
}
extern "C" int nabu_speller_bwd(const nabu_speller_desc_t*, const nabu_speller_params_t*, const float*, const int*,
                                const int*, int, const int*, const float*, void*, float*,
                                const nabu_speller_params_t*, void*, size_t, void*) {
  NOT_BUILT("nabu_speller_bwd");
}
extern "C" size_t nabu_las_beam_workspace_bytes(const nabu_speller_desc_t*, int, int) { return 0; }
extern "C" int nabu_las_beam_search(const nabu_speller_desc_t*, const nabu_speller_params_t*, const float*,
                                    const int*, int, int, float, float, int*, int*, float*, float*, int*, void*,
                                    size_t, void*) {
  NOT_BUILT("nabu_las_beam_search");
}
extern "C" size_t nabu_ctc_beam_workspace_bytes(int, int, int, int) { return 0; }
extern "C" int nabu_ctc_beam_search(const float*, const int*, int, int, int, int, int, int*, int*, float*, void*,
                                    size_t, void*) {
  NOT_BUILT("nabu_ctc_beam_search");
}
