// fmt_harness.cpp -- TEST INFRASTRUCTURE.  Links the product's host-side printers (popbam_b200/csrc/pb_format.cpp)
// without the CUDA library, so that their text can be checked on a machine without a GPU: the test feeds them the
// oracle's result struct (same layout as pb_region_result) and compares with the reference goldens.
#include "../include/popbam_b200.h"

static pb_params g_params;
extern "C" void fmt_set_params(const pb_params *p) { g_params = *p; }
// the one symbol pb_format.cpp takes from pb_lib.cu
extern "C" const pb_params *pb_ctx_params(const pb_ctx *) { return &g_params; }
