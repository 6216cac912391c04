// gemm.cuh -- the FP64 tensor-core (DMMA) GEMM every O(N^3) step is built from.
#pragma once

#include "common.cuh"

namespace pgp {

// C (M x N, row-major, ldc) = beta * C + alpha * A (M x K) * B (N x K)^T
// A and B are row-major with the contraction index contiguous ("NT").
struct GemmArgs {
    const double* A = nullptr;
    const double* B = nullptr;
    double* C = nullptr;
    int64_t lda = 0, ldb = 0, ldc = 0;
    int64_t M = 0, N = 0, K = 0;
    double alpha = 1.0, beta = 0.0;
    // tri: skip C tiles lying strictly above the diagonal row - col = tri_off
    // (entry (i, j) is "on or below" when j <= i + tri_off).
    int tri = 0;
    int64_t tri_off = 0;
    // krow: A is upper-triangular in (row, k): A[i][k] == 0 for k < i + krow_off,
    // so the contraction of tile row i0 may start at k = i0 + krow_off
    // (rounded down to the k-tile).  Used by the triangular inverse and by
    // V V^T (k >= max(i, j) = i on the lower tiles).
    int krow = 0;
    int64_t krow_off = 0;
    // kcol (with transB): B is upper-triangular in (k, col): B[k][j] == 0 for
    // k > j + kcol_off, so the contraction of tile column j0 may stop at
    // k = j0 + BN + kcol_off.  Used by the bottom-up triangular inverse.
    int kcol = 0;
    int64_t kcol_off = 0;
    // stair: the rows of A and C are `stair_front` dense rows followed by blocks of stair_nb rows, block q
    // of which only exists from column S(q) = stair_first + q stair_step of the caller's column space
    // (chol.cuh: Stair).  Bit 1: A's block q is zero for k + stair_off < S(q), so the contraction of a tile
    // row may start there (as krow, with a slope of stair_step / stair_nb).  Bit 2: C's block q is only
    // needed at columns n + stair_off >= S(q): tiles wholly to the left are skipped.
    int stair = 0;
    int64_t stair_front = 0, stair_nb = 1, stair_first = 0, stair_step = 0, stair_off = 0;
    int batch = 1;
    int64_t strideA = 0, strideB = 0, strideC = 0;
    // ragged batch: member b has M - b * ragged_mstep rows (the block columns of one trailing update of the
    // distributed factorisation: equal width, each starting ragged_mstep rows further down)
    int64_t ragged_mstep = 0;
    // general forms (launch_gemm): transA -> A is stored (K x M) row-major,
    // transB -> B is stored (K x N) row-major, i.e.
    //   NT (0,0): C = A B^T     NN (0,1): C = A B     TN (1,1): C = A^T B
    int transA = 0, transB = 0;
    // contraction split: 0 = automatic (few output tiles, long K), 1 = off,
    // > 1 = that many slices.  Partials go to the context workspace and are
    // reduced in a fixed order (deterministic).
    int splitk = 1;
    double* ws = nullptr;   // set by the launcher
    int64_t ldws = 0;
};

// NT form, no split (the factorisation's work-horse)
int launch_gemm_nt(pgp_ctx* ctx, const GemmArgs& a);
// any form; honours transA / transB / splitk
int launch_gemm(pgp_ctx* ctx, const GemmArgs& a);

}  // namespace pgp
