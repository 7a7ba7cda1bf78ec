// ref_shim.cu -- extern "C" doorway onto the UNMODIFIED reference host utilities.
// TEST INFRASTRUCTURE ONLY (see oracle/oracle.c header).  Compiled by oracle/Makefile
// together with /root/reference/utils/utils.cu (read in place, never copied) into
// oracle/_ref/libref_utils.so, which tests use to pin oracle.c bit-for-bit.
#include <cstdlib>
#include "utils.cuh"  // resolved via -I/root/reference/utils

extern "C" {
void ref_srand(unsigned int s) { srand(s); }
void ref_generate_random_matrix(float *t, int n) { generate_random_matrix(t, n); }
void ref_fill_vector(float *t, float v, int n) { fill_vector(t, v, n); }
void ref_cpu_gemm(float alpha, float beta, float *x, float *y, int n, float *z) {
  cpu_gemm(alpha, beta, x, y, n, z);
}
int ref_verify_matrix(float *a, float *b, int m, int n) { return verify_matrix(a, b, m, n) ? 1 : 0; }
}
