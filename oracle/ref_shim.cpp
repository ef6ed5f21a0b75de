/*
 * oracle/ref_shim.cpp -- TEST INFRASTRUCTURE.  extern "C" re-exports of the reference's own CPU checker
 * functions (matInit, stackInit, stackCalc, stackTransp, checkSum, checkSumTransp) so that ctypes can call
 * them.  The function bodies are NOT here: they are compiled from the reference source where it lies
 * (/root/reference/src/acc/libsmm_acc/libsmm_acc_benchmark.cpp) by oracle/Makefile into oracle/_ref/.
 * `ht` is the autotuned-parameter table the reference's benchmark() looks up (generated parameters.h in a
 * real build); an empty table satisfies the linker, benchmark() itself is never called from here.
 */
#include "libsmm_acc_benchmark.h"
#include "parameters_utils.h"

extern const std::unordered_map<Triplet, KernelParameters> ht = {};
/* flag objects declared in the reference's src/acc/cuda/acc_cuda.h:101-102 (defined in acc_cuda.cpp, which also
 * drags in NVRTC/cuBLAS and is therefore not compiled here) */
CUevent_flags CUEventDefault = CU_EVENT_DEFAULT;
CUstream_flags CUStreamDefault = CU_STREAM_DEFAULT;

extern "C" {
void ref_matInit(double* mat, int mat_n, int x, int y, int seed) { matInit(mat, mat_n, x, y, seed); }
void ref_stackInit(int* stack, int n_stack, int n_c, int n_a, int n_b, int m, int n, int k) {
  stackInit(stack, n_stack, n_c, n_a, n_b, m, n, k);
}
void ref_stackCalc(int* stack, int n_stack, double* c, double* a, double* b, int m, int n, int k) {
  stackCalc(stack, n_stack, c, a, b, m, n, k);
}
void ref_stackTransp(int* stack, int n_stack, double* mat, double* mat_trs, int m, int n) {
  stackTransp(stack, n_stack, mat, mat_trs, m, n);
}
double ref_checkSum(double* c, int n_c, int m, int n) { return checkSum(c, n_c, m, n); }
double ref_checkSumTransp(double* mat, int n_stack, int m, int n) { return checkSumTransp(mat, n_stack, m, n); }
}
