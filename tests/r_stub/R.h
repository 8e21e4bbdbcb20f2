/* Minimal DECLARATIONS of the part of R's public C API that r-package/src/bpcuda_rcall.c uses, written for this repository's
 * tests only (tests/test_cabi.py::test_r_shim_compiles_against_the_header): R is not installed in the build image, and a syntax check
 * of the shim against include/bpcuda.h still catches a wrong argument list or a missing symbol.  Not R's headers; nothing here is
 * linked or run. */
#ifndef BPC_TEST_R_STUB_H
#define BPC_TEST_R_STUB_H
#include <stddef.h>
char* R_alloc(size_t n, int size);
#endif
