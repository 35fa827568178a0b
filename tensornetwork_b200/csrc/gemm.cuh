// gemm.cuh — the canonical (batched, strided) GEMM problem the tensordot planner lowers to.
#pragma once
#include "common.cuh"

namespace tnb {

// C[b, m, n] = sum_k opA(A[b, m, k]) * opB(B[b, k, n]); all strides in elements.
struct GemmProblem {
  int dtype = 0;
  int64_t M = 0, N = 0, K = 0, batch = 1;
  const void* A = nullptr; int64_t a_sm = 0, a_sk = 0, a_sb = 0;
  const void* B = nullptr; int64_t b_sk = 0, b_sn = 0, b_sb = 0;
  void* C = nullptr;       int64_t c_sm = 0, c_sn = 0, c_sb = 0;
  bool conjA = false, conjB = false;
  int math = 0;  // TNB200_MATH_* >> 4
};

// Each returns TNB200_ERR_UNSUPPORTED (without setting an error) when the problem does not
// meet the kernel's layout/alignment constraints; the planner then repacks or falls back.
int gemm_tcgen05(const GemmProblem& p, cudaStream_t st);   // bf16 / f16 / f32(tf32)
int gemm_dmma_f64(const GemmProblem& p, cudaStream_t st);  // f64 via mma.sync DMMA
bool tcgen05_operand_ok(int dtype, const void* ptr, int64_t ext_mn, int64_t ext_k,
                        int64_t s_mn, int64_t s_k, int64_t s_b, int64_t batch);

}  // namespace tnb
