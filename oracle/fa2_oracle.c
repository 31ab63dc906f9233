/*
 * fa2_oracle.c -- CPU restatement of the reference's FlashAttention-2 forward and
 * backward algorithm (detker/CUDA-Flash-Attention).
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing in the product path (libfa2_b200.so, the CLI,
 * the fa2_b200 Python package) may call, link or import this file.  Allowed users:
 * tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs,
 * and there only as the checker / reported baseline.
 *
 * Parity pin: this restatement is checked (tests/test_oracle.py) against golden
 * vectors produced by the reference's own Python oracle
 * (test_flash_attention2.py:197-208 compute_reference, :220-232 autograd grads,
 * :917-921 LSE) -- see tests/golden/make_golden.py -- and, on the GPU box, against
 * the unmodified reference CUDA kernels compiled into oracle/_ref/ (oracle/build_ref.sh).
 *
 * What it follows (all paths relative to /root/reference):
 *   forward   kernels/kernel_fa2_optimized.cu:19-347   (32x32 tiles, online softmax)
 *   D pre-pass kernels/f-attn2-backward.cu:342-380
 *   backward  kernels/f-attn2-backward.cu:33-339       (one 32-row KV tile vs all Q tiles)
 * Arithmetic is float32 in the reference's operation order; the fast-math device
 * intrinsics __expf/__logf are restated with expf/logf (a few ulp apart).
 * Layout: [B,H,S,D] row-major contiguous; LSE and D_i are [B,H,S].
 *
 * Build: gcc -O2 -fopenmp -shared -fPIC (see oracle/Makefile).
 */
#include <float.h>
#include <math.h>
#include <stddef.h>
#include <stdlib.h>
#include <string.h>

#define TILE_R 32 /* BLOCK_SIZE_R, kernel_fa2_optimized.cu:388-389 */
#define TILE_C 32 /* BLOCK_SIZE_C */

int fa2_oracle_version(void) { return 1; }

/*
 * Forward: O = softmax(Q K^T / sqrt(D)) V,  LSE = ln(sum exp) + max  (natural log).
 * One (b,h,q-tile) "block" at a time, exactly the loop nest of
 * kernel_fa2_optimized.cu:89-325 with the epilogue of :327-346.
 */
int fa2_oracle_forward(const float *Q, const float *K, const float *V, float *O, float *LSE,
                       int B, int H, int S, int D)
{
    if (B <= 0 || H <= 0 || S <= 0 || D <= 0) return 1;
    const float sqrt_d = sqrtf((float)D); /* :51 */
    const int T_r = (S + TILE_R - 1) / TILE_R;
    const int T_c = (S + TILE_C - 1) / TILE_C;
    const long n_blocks = (long)B * H * T_r;
    int failed = 0;

#pragma omp parallel for schedule(dynamic, 1)
    for (long blk = 0; blk < n_blocks; ++blk) {
        const long bh = blk / T_r;
        const int qt = (int)(blk % T_r);
        const float *q = Q + (size_t)bh * S * D;
        const float *k = K + (size_t)bh * S * D;
        const float *v = V + (size_t)bh * S * D;
        float *o = O + (size_t)bh * S * D;
        float *lse = LSE + (size_t)bh * S;

        float *o_acc = (float *)calloc((size_t)TILE_R * D, sizeof(float));
        if (!o_acc) { failed = 1; continue; }
        float s_buf[TILE_R][TILE_C];
        float l_run[TILE_R], m_run[TILE_R], coeff[TILE_R];
        for (int r = 0; r < TILE_R; ++r) { l_run[r] = 0.0f; m_run[r] = -FLT_MAX; } /* :77-83 */

        for (int j = 0; j < T_c; ++j) {
            /* S = Q K^T / sqrt(D), padded columns = -FLT_MAX   (:126-191) */
            for (int r = 0; r < TILE_R; ++r) {
                const int qi = qt * TILE_R + r;
                if (qi >= S) continue;
                for (int c = 0; c < TILE_C; ++c) {
                    const int kj = j * TILE_C + c;
                    if (kj >= S) { s_buf[r][c] = -FLT_MAX; continue; }
                    float acc = 0.0f;
                    for (int d = 0; d < D; ++d) acc += q[(size_t)qi * D + d] * k[(size_t)kj * D + d];
                    s_buf[r][c] = acc / sqrt_d; /* division, :187 */
                }
            }
            /* online softmax per row   (:194-255) */
            for (int r = 0; r < TILE_R; ++r) {
                const int qi = qt * TILE_R + r;
                if (qi >= S) continue;
                float row_max = -FLT_MAX;
                for (int c = 0; c < TILE_C; ++c) row_max = fmaxf(row_max, s_buf[r][c]);
                const float new_max = fmaxf(m_run[r], row_max);
                coeff[r] = expf(m_run[r] - new_max);
                m_run[r] = new_max;
                float row_sum = 0.0f;
                for (int c = 0; c < TILE_C; ++c) {
                    s_buf[r][c] = expf(s_buf[r][c] - new_max);
                    row_sum += s_buf[r][c];
                }
                l_run[r] = coeff[r] * l_run[r] + row_sum;
            }
            /* O = O * coeff + P V   (:286-324); padded V rows are zero */
            for (int r = 0; r < TILE_R; ++r) {
                const int qi = qt * TILE_R + r;
                if (qi >= S) continue;
                for (int d = 0; d < D; ++d) {
                    float acc = 0.0f;
                    for (int c = 0; c < TILE_C; ++c) {
                        const int kj = j * TILE_C + c;
                        const float vv = (kj < S) ? v[(size_t)kj * D + d] : 0.0f;
                        acc += s_buf[r][c] * vv;
                    }
                    o_acc[(size_t)r * D + d] = o_acc[(size_t)r * D + d] * coeff[r] + acc;
                }
            }
        }
        /* epilogue: O /= l ; LSE = ln(l) + m   (:327-346) */
        for (int r = 0; r < TILE_R; ++r) {
            const int qi = qt * TILE_R + r;
            if (qi >= S) continue;
            for (int d = 0; d < D; ++d) o[(size_t)qi * D + d] = o_acc[(size_t)r * D + d] / l_run[r];
            lse[qi] = logf(l_run[r]) + m_run[r];
        }
        free(o_acc);
    }
    return failed;
}

/* D_i = sum_d dO[i,d] * O[i,d]   (f-attn2-backward.cu:342-380), N = B*H*S rows. */
int fa2_oracle_rowdot(const float *dO, const float *O, float *Dvec, long N, int D)
{
    if (N < 0 || D <= 0) return 1;
#pragma omp parallel for schedule(static)
    for (long i = 0; i < N; ++i) {
        float acc = 0.0f;
        for (int d = 0; d < D; ++d) acc += dO[(size_t)i * D + d] * O[(size_t)i * D + d];
        Dvec[i] = acc;
    }
    return 0;
}

/*
 * Backward (f-attn2-backward.cu:33-339).  One 32-row KV tile per "block"; for every
 * 32-row Q tile: P = exp(QK^T/sqrt(D) - LSE) (:152-183), dV += P^T dO (:219-239),
 * dP = dO V^T, dS = P*(dP - D_i)/sqrt(D) (:243-266), dQ += dS K (:270-300, global
 * atomicAdd in the reference), dK += dS^T Q (:304-322).
 * dQ is accumulated per (b,h) serially over KV tiles here, so the result is
 * deterministic (the reference's atomics are not).
 */
int fa2_oracle_backward(const float *Q, const float *K, const float *V, const float *O,
                        const float *dO, const float *LSE, float *dQ, float *dK, float *dV,
                        int B, int H, int S, int D)
{
    if (B <= 0 || H <= 0 || S <= 0 || D <= 0) return 1;
    const float sqrt_d = sqrtf((float)D);
    const int T_r = (S + TILE_R - 1) / TILE_R;
    const int T_c = (S + TILE_C - 1) / TILE_C;
    const long BH = (long)B * H;
    int failed = 0;

#pragma omp parallel for schedule(dynamic, 1)
    for (long bh = 0; bh < BH; ++bh) {
        const float *q = Q + (size_t)bh * S * D;
        const float *k = K + (size_t)bh * S * D;
        const float *v = V + (size_t)bh * S * D;
        const float *o = O + (size_t)bh * S * D;
        const float *go = dO + (size_t)bh * S * D;
        const float *lse = LSE + (size_t)bh * S;
        float *gq = dQ + (size_t)bh * S * D;
        float *gk = dK + (size_t)bh * S * D;
        float *gv = dV + (size_t)bh * S * D;

        float *Di = (float *)malloc((size_t)S * sizeof(float));
        float *dk_acc = (float *)malloc((size_t)TILE_C * D * sizeof(float));
        float *dv_acc = (float *)malloc((size_t)TILE_C * D * sizeof(float));
        if (!Di || !dk_acc || !dv_acc) { failed = 1; free(Di); free(dk_acc); free(dv_acc); continue; }
        for (int i = 0; i < S; ++i) {
            float acc = 0.0f;
            for (int d = 0; d < D; ++d) acc += go[(size_t)i * D + d] * o[(size_t)i * D + d];
            Di[i] = acc;
        }
        memset(gq, 0, (size_t)S * D * sizeof(float)); /* cudaMemset, f-attn2-backward.cu:427 */

        float p_buf[TILE_R][TILE_C];
        for (int jt = 0; jt < T_c; ++jt) {
            memset(dk_acc, 0, (size_t)TILE_C * D * sizeof(float));
            memset(dv_acc, 0, (size_t)TILE_C * D * sizeof(float));
            for (int it = 0; it < T_r; ++it) {
                /* P_ij */
                for (int r = 0; r < TILE_R; ++r) {
                    const int qi = it * TILE_R + r;
                    for (int c = 0; c < TILE_C; ++c) {
                        const int kj = jt * TILE_C + c;
                        if (qi >= S || kj >= S) { p_buf[r][c] = 0.0f; continue; }
                        float acc = 0.0f;
                        for (int d = 0; d < D; ++d) acc = fmaf(q[(size_t)qi * D + d], k[(size_t)kj * D + d], acc);
                        acc /= sqrt_d;
                        p_buf[r][c] = expf(acc - lse[qi]);
                    }
                }
                /* dV_j += P^T dO_i */
                for (int c = 0; c < TILE_C; ++c) {
                    const int kj = jt * TILE_C + c;
                    if (kj >= S) continue;
                    for (int d = 0; d < D; ++d) {
                        float acc = 0.0f;
                        for (int r = 0; r < TILE_R; ++r) {
                            const int qi = it * TILE_R + r;
                            if (qi < S) acc += p_buf[r][c] * go[(size_t)qi * D + d];
                        }
                        dv_acc[(size_t)c * D + d] += acc;
                    }
                }
                /* dS = P * (dP - D_i) / sqrt(D), dP = dO V^T ; stored over P */
                for (int r = 0; r < TILE_R; ++r) {
                    const int qi = it * TILE_R + r;
                    if (qi >= S) continue;
                    for (int c = 0; c < TILE_C; ++c) {
                        const int kj = jt * TILE_C + c;
                        if (kj >= S) continue;
                        float dp = 0.0f;
                        for (int d = 0; d < D; ++d) dp += go[(size_t)qi * D + d] * v[(size_t)kj * D + d];
                        p_buf[r][c] = (dp - Di[qi]) * p_buf[r][c] / sqrt_d;
                    }
                }
                /* dQ_i += dS K_j */
                for (int r = 0; r < TILE_R; ++r) {
                    const int qi = it * TILE_R + r;
                    if (qi >= S) continue;
                    for (int d = 0; d < D; ++d) {
                        float acc = 0.0f;
                        for (int c = 0; c < TILE_C; ++c) {
                            const int kj = jt * TILE_C + c;
                            if (kj < S) acc += p_buf[r][c] * k[(size_t)kj * D + d];
                        }
                        gq[(size_t)qi * D + d] += acc;
                    }
                }
                /* dK_j += dS^T Q_i */
                for (int c = 0; c < TILE_C; ++c) {
                    const int kj = jt * TILE_C + c;
                    if (kj >= S) continue;
                    for (int d = 0; d < D; ++d) {
                        float acc = 0.0f;
                        for (int r = 0; r < TILE_R; ++r) {
                            const int qi = it * TILE_R + r;
                            if (qi < S) acc += p_buf[r][c] * q[(size_t)qi * D + d];
                        }
                        dk_acc[(size_t)c * D + d] += acc;
                    }
                }
            }
            for (int c = 0; c < TILE_C; ++c) {
                const int kj = jt * TILE_C + c;
                if (kj >= S) continue;
                memcpy(gk + (size_t)kj * D, dk_acc + (size_t)c * D, (size_t)D * sizeof(float));
                memcpy(gv + (size_t)kj * D, dv_acc + (size_t)c * D, (size_t)D * sizeof(float));
            }
        }
        free(Di); free(dk_acc); free(dv_acc);
    }
    return failed;
}
