// Host shim around the matrix-derived code classes of segalign_b200/csrc/screen_bound.h (tests/test_code_classes.py).
#include <cstdint>
#include "../../segalign_b200/csrc/screen_bound.h"
extern "C" void sa_test_code_classes(const int *sub_mat, int xdrop, uint32_t *out) {
    out[0] = sa::screen_terminator_codes(sub_mat, xdrop, &out[1]);
    sa::zero_run_codes(sub_mat, &out[2], &out[3]);
}
