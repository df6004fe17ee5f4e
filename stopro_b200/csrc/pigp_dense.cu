// FP64 dense linear algebra for the PIGP path on sm_100a (K3/K4/K5/K7).
//
// Replaces the library calls of the reference: jnp.linalg.cholesky (GP/gp.py:83, :106, :430), the general
// jnp.linalg.solve calls on the triangular factor (:84, :109, :118, :432-433) and the per-parameter
// Sigma_inv @ dK matmuls (:481-485).
//
// * k_gemm: C = alpha * A * B^T + beta * C on 128 x 128 tiles, operands staged through shared memory by a
//   3-stage cp.async pipeline, the products issued as FP64 tensor-core MMAs (mma.sync.m8n8k4.f64, SASS DMMA.8x8x4;
//   tcgen05 has no f64 kind).  Either index of A / B may be the contiguous one, k-ranges can follow a triangular
//   operand tile by tile and tiles above the diagonal can be skipped, so that the same kernel is the SYRK/GEMM
//   trailing update of the Cholesky, the TRSM-by-inverse, both TRTRI products and the LAUUM.
// * k_potf2: one CTA factors a 128 x 128 diagonal tile in shared memory and inverts the factor in place.
// * Host drivers: recursive right-looking Cholesky (all flops in k_gemm), TRTRI + LAUUM for K^-1.
#include <algorithm>

#include "pigp_internal.cuh"

namespace pigp {

// ----------------------------------------------------------------------------------------------- GEMM
constexpr int BM = 128, BN = 128, BK = 16, STAGES = 3, GEMM_THREADS = 256;
constexpr int LDS_K = BK + 4;   // k-contiguous operand tile  [128][20]  (row stride = 4 mod 16 doubles: conflict-free fragment loads)
constexpr int LDS_M = BM + 4;   // m-contiguous operand tile  [16][132]
constexpr int OPD = BM * LDS_K; // doubles per operand per stage (2560 >= 16*132)
constexpr int GEMM_SMEM = STAGES * 2 * OPD * (int)sizeof(double);

__device__ __forceinline__ void cp_async16(void* smem, const void* gmem) {
    const unsigned s = (unsigned)__cvta_generic_to_shared(smem);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(s), "l"(gmem));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
    asm volatile("cp.async.wait_group %0;\n" ::"n"(N));
}
__device__ __forceinline__ void dmma(double& c0, double& c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                 : "+d"(c0), "+d"(c1)
                 : "d"(a), "d"(b));
}

template <bool KC>
__device__ __forceinline__ void load_operand(double* sm, const double* G, int64_t ld, int64_t row0, int64_t k0, int tid) {
    if (KC) {  // X(row, k) = G[row*ld + k]: 128 rows x 16 doubles, 8 16-byte chunks per row
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int c = tid + i * GEMM_THREADS;
            const int r = c >> 3, ch = c & 7;
            cp_async16(sm + r * LDS_K + ch * 2, G + (row0 + r) * ld + k0 + ch * 2);
        }
    } else {   // X(row, k) = G[k*ld + row]: 16 k-rows x 128 doubles, 64 chunks per k-row
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int c = tid + i * GEMM_THREADS;
            const int kr = c >> 6, ch = c & 63;
            cp_async16(sm + kr * LDS_M + ch * 2, G + (k0 + kr) * ld + row0 + ch * 2);
        }
    }
}

template <bool AKC, bool BKC>
__global__ void __launch_bounds__(GEMM_THREADS, 1) k_gemm(GemmDesc g) {
    extern __shared__ __align__(16) double smem[];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int gid = lane >> 2, tig = lane & 3;
    const int wm = warp & 1, wn = warp >> 1;  // 2 x 4 warps, warp tile 64 x 32

    int tm, tn;
    if (g.lower_only) {
        const int idx = blockIdx.x;
        tm = (int)((sqrt(8.0 * idx + 1.0) - 1.0) * 0.5);
        while ((tm + 1) * (tm + 2) / 2 <= idx) ++tm;
        while (tm * (tm + 1) / 2 > idx) --tm;
        tn = idx - tm * (tm + 1) / 2;
    } else {
        const int mt = g.M / BM;
        tm = blockIdx.x % mt;
        tn = blockIdx.x / mt;
    }
    const int64_t m0 = (int64_t)tm * BM, n0 = (int64_t)tn * BN;
    int kt_begin = 0, kt_end = g.K / BK;
    if (g.kmode == 1) kt_begin = tm * (BM / BK);
    else if (g.kmode == 2) kt_end = min(kt_end, (tm + 1) * (BM / BK));

    double acc[8][4][2];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;

    const int nk = kt_end - kt_begin;
#pragma unroll
    for (int s = 0; s < STAGES - 1; ++s) {
        if (s < nk) {
            load_operand<AKC>(smem + s * 2 * OPD, g.A, g.lda, m0, (int64_t)(kt_begin + s) * BK, tid);
            load_operand<BKC>(smem + s * 2 * OPD + OPD, g.B, g.ldb, n0, (int64_t)(kt_begin + s) * BK, tid);
        }
        cp_async_commit();
    }
    for (int it = 0; it < nk; ++it) {
        cp_async_wait<STAGES - 2>();
        __syncthreads();
        {
            const int nx = it + STAGES - 1;
            if (nx < nk) {
                const int s = nx % STAGES;
                load_operand<AKC>(smem + s * 2 * OPD, g.A, g.lda, m0, (int64_t)(kt_begin + nx) * BK, tid);
                load_operand<BKC>(smem + s * 2 * OPD + OPD, g.B, g.ldb, n0, (int64_t)(kt_begin + nx) * BK, tid);
            }
            cp_async_commit();
        }
        const double* sA = smem + (it % STAGES) * 2 * OPD;
        const double* sB = sA + OPD;
#pragma unroll
        for (int kk = 0; kk < BK / 4; ++kk) {
            double af[8], bf[4];
#pragma unroll
            for (int mi = 0; mi < 8; ++mi) {
                const int r = wm * 64 + mi * 8 + gid, k = kk * 4 + tig;
                af[mi] = AKC ? sA[r * LDS_K + k] : sA[k * LDS_M + r];
            }
#pragma unroll
            for (int ni = 0; ni < 4; ++ni) {
                const int r = wn * 32 + ni * 8 + gid, k = kk * 4 + tig;
                bf[ni] = BKC ? sB[r * LDS_K + k] : sB[k * LDS_M + r];
            }
#pragma unroll
            for (int mi = 0; mi < 8; ++mi)
#pragma unroll
                for (int ni = 0; ni < 4; ++ni) dmma(acc[mi][ni][0], acc[mi][ni][1], af[mi], bf[ni]);
        }
    }
    cp_async_wait<0>();

#pragma unroll
    for (int mi = 0; mi < 8; ++mi) {
        const int64_t r = m0 + wm * 64 + mi * 8 + gid;
#pragma unroll
        for (int ni = 0; ni < 4; ++ni) {
            const int64_t c = n0 + wn * 32 + ni * 8 + tig * 2;
            double2* p = reinterpret_cast<double2*>(g.C + r * g.ldc + c);
            double2 o;
            o.x = g.alpha * acc[mi][ni][0];
            o.y = g.alpha * acc[mi][ni][1];
            if (g.beta != 0.0) {
                const double2 old = *p;
                o.x = fma(g.beta, old.x, o.x);
                o.y = fma(g.beta, old.y, o.y);
            }
            *p = o;
        }
    }
}

int launch_gemm(const GemmDesc& g, cudaStream_t st) {
    if (g.M <= 0 || g.N <= 0) return PIGP_OK;
    if (g.M % BM || g.N % BN || g.K % BK || g.K <= 0) {
        set_error("pigp gemm: M, N must be multiples of 128 and K of 16");
        return PIGP_EINVAL;
    }
    const int mt = g.M / BM, nt = g.N / BN;
    int64_t tiles;
    if (g.lower_only) {
        // tiles (tm, tn <= tm); rows beyond the square part (tm >= nt) have all nt column tiles
        if (mt < nt) { set_error("pigp gemm: lower_only needs M >= N"); return PIGP_EINVAL; }
        tiles = (int64_t)nt * (nt + 1) / 2;
    } else {
        tiles = (int64_t)mt * nt;
    }
    static bool attr_done[64] = {};
    int dev = 0;
    PIGP_CUDA(cudaGetDevice(&dev));
    bool& attr_set = attr_done[dev & 63];
    auto set_attr = [&](auto kern) { return cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, GEMM_SMEM); };
    if (!attr_set) {
        PIGP_CUDA(set_attr(k_gemm<true, true>));
        PIGP_CUDA(set_attr(k_gemm<true, false>));
        PIGP_CUDA(set_attr(k_gemm<false, true>));
        PIGP_CUDA(set_attr(k_gemm<false, false>));
        attr_set = true;
    }
    auto run = [&](const GemmDesc& d, int64_t ntiles) {
        const dim3 grid((unsigned)ntiles), block(GEMM_THREADS);
        double flops = 0.0;
        if (g_prof_on) {  // flops executed at tile granularity
            const int kt = d.K / BK, per = BM / BK;
            const int dnt = d.N / BN;
            const int dmt = d.lower_only ? dnt : d.M / BM;  // the lower_only launch covers the square part only
            for (int tm = 0; tm < dmt; ++tm) {
                const int ncols = d.lower_only ? tm + 1 : dnt;
                int kb = 0, ke = kt;
                if (d.kmode == 1) kb = tm * per;
                else if (d.kmode == 2) ke = std::min(kt, (tm + 1) * per);
                flops += 2.0 * BM * BN * BK * (double)std::max(0, ke - kb) * ncols;
            }
            prof_note(d.lower_only ? d.N : d.M, d.N, d.K, d.kmode * 10 + d.lower_only);
        }
        ProfScope prof(PROF_GEMM, st, flops);
        if (d.a_kcontig && d.b_kcontig) k_gemm<true, true><<<grid, block, GEMM_SMEM, st>>>(d);
        else if (d.a_kcontig) k_gemm<true, false><<<grid, block, GEMM_SMEM, st>>>(d);
        else if (d.b_kcontig) k_gemm<false, true><<<grid, block, GEMM_SMEM, st>>>(d);
        else k_gemm<false, false><<<grid, block, GEMM_SMEM, st>>>(d);
        count_launch();
    };
    run(g, tiles);
    if (g.lower_only && mt > nt) {
        // rectangular remainder below the square part: rows [N, M)
        GemmDesc r = g;
        r.lower_only = 0;
        r.M = g.M - g.N;
        r.A = g.a_kcontig ? g.A + (int64_t)g.N * g.lda : g.A + g.N;
        r.C = g.C + (int64_t)g.N * g.ldc;
        if (g.kmode != 0) { set_error("pigp gemm: kmode with rectangular lower_only is unsupported"); return PIGP_EINVAL; }
        run(r, (int64_t)(r.M / BM) * nt);
    }
    PIGP_CUDA(cudaGetLastError());
    return PIGP_OK;
}

// ----------------------------------------------------------------------------------------------- POTF2 (128 x 128)
constexpr int PT = 128;
constexpr int PLD = PT + 1;  // odd row stride: column walks are conflict-free
constexpr int POTF2_SMEM = PT * PLD * (int)sizeof(double);

// Factor the lower triangle of the 128 x 128 tile at A in place (upper part of the tile is set to zero) and write
// inv(L) (lower, zeros above) to invd[128*128].  Non-positive pivot -> *info = base + column + 1 (first one wins)
// and NaNs propagate, which is what jnp.linalg.cholesky gives the reference.
__global__ void __launch_bounds__(256, 1) k_potf2(double* A, int64_t ld, double* invd, int32_t* info, int base) {
    extern __shared__ __align__(16) double sm[];
    __shared__ double s_part[2][PT];
    const int tid = threadIdx.x;
    for (int e = tid; e < PT * PT; e += 256) {
        const int i = e >> 7, j = e & 127;
        sm[i * PLD + j] = (j <= i) ? A[(int64_t)i * ld + j] : 0.0;
    }
    const int row = tid & 127, half = tid >> 7;
    for (int j = 0; j < PT; ++j) {
        __syncthreads();
        const double d = sm[j * PLD + j];
        if (tid == 0 && !(d > 0.0) && info) atomicCAS(info, 0, base + j + 1);
        const double rinv = 1.0 / sqrt(d);
        __syncthreads();
        // scale column j
        if (half == 0) {
            if (row > j) sm[row * PLD + j] *= rinv;
            else if (row == j) sm[j * PLD + j] = d * rinv;
        }
        __syncthreads();
        // trailing update: row `row`, columns k in (j, row], interleaved over the two halves
        if (row > j) {
            const double lij = sm[row * PLD + j];
            double* mine = sm + row * PLD;
            int k = j + 1 + half;
            for (; k + 6 <= row; k += 8) {  // four independent updates per trip: loads first, stores last
                const double l0 = sm[k * PLD + j], l1 = sm[(k + 2) * PLD + j], l2 = sm[(k + 4) * PLD + j], l3 = sm[(k + 6) * PLD + j];
                const double c0 = mine[k], c1 = mine[k + 2], c2 = mine[k + 4], c3 = mine[k + 6];
                mine[k] = fma(-lij, l0, c0);
                mine[k + 2] = fma(-lij, l1, c1);
                mine[k + 4] = fma(-lij, l2, c2);
                mine[k + 6] = fma(-lij, l3, c3);
            }
            for (; k <= row; k += 2) mine[k] = fma(-lij, sm[k * PLD + j], mine[k]);
        }
    }
    __syncthreads();
    for (int e = tid; e < PT * PT; e += 256) {
        const int i = e >> 7, j = e & 127;
        A[(int64_t)i * ld + j] = sm[i * PLD + j];
    }
    // in-place inverse, row by row: W[i][j] = (delta_ij - sum_{k=j}^{i-1} L[i][k] W[k][j]) / L[i][i]
    // thread (col = row, half): half 0 takes even offsets of k, half 1 odd offsets
    const int col = row;
    for (int i = 0; i < PT; ++i) {
        __syncthreads();
        double s0 = 0.0, s1 = 0.0;
        if (col < i) {
            int k = col + half;
            for (; k + 2 < i; k += 4) {
                s0 = fma(sm[i * PLD + k], sm[k * PLD + col], s0);
                s1 = fma(sm[i * PLD + k + 2], sm[(k + 2) * PLD + col], s1);
            }
            for (; k < i; k += 2) s0 = fma(sm[i * PLD + k], sm[k * PLD + col], s0);
        }
        s_part[half][col] = s0 + s1;
        const double dii = sm[i * PLD + i];
        __syncthreads();
        if (half == 0 && col <= i) {
            const double rhs = (col == i) ? 1.0 : 0.0;
            sm[i * PLD + col] = (rhs - (s_part[0][col] + s_part[1][col])) / dii;
        }
    }
    __syncthreads();
    for (int e = tid; e < PT * PT; e += 256) {
        const int i = e >> 7, j = e & 127;
        invd[e] = sm[i * PLD + j];
    }
}

static int launch_potf2(double* A, int64_t ld, double* invd, int32_t* info, int base, cudaStream_t st) {
    static bool attr_done[64] = {};
    int dev = 0;
    PIGP_CUDA(cudaGetDevice(&dev));
    bool& attr_set = attr_done[dev & 63];
    if (!attr_set) {
        PIGP_CUDA(cudaFuncSetAttribute(k_potf2, cudaFuncAttributeMaxDynamicSharedMemorySize, POTF2_SMEM));
        attr_set = true;
    }
    ProfScope prof(PROF_POTF2, st);
    k_potf2<<<1, 256, POTF2_SMEM, st>>>(A, ld, invd, info, base);
    count_launch();
    PIGP_CUDA(cudaGetLastError());
    return PIGP_OK;
}

// Factor A[0:n,0:n] and apply L^-T from the right to the m_below rows under it (recursive right-looking).
static int chol_rec(double* A, int64_t ld, int64_t n, int64_t m_below, double* invd, int32_t* info, int base,
                    cudaStream_t st) {
    if (n == TILE) {
        PIGP_TRY(launch_potf2(A, ld, invd, info, base, st));
        if (m_below > 0) {
            // B <- B * inv(L)^T, in place: every CTA reads its own 128 rows completely before writing them
            GemmDesc g{};
            g.M = (int)m_below; g.N = TILE; g.K = TILE;
            g.alpha = 1.0; g.beta = 0.0;
            g.A = A + TILE * ld; g.lda = ld; g.a_kcontig = 1;
            g.B = invd; g.ldb = TILE; g.b_kcontig = 1;
            g.C = A + TILE * ld; g.ldc = ld;
            PIGP_TRY(launch_gemm(g, st));
        }
        return PIGP_OK;
    }
    const int64_t n1 = (n / TILE / 2) * TILE, n2 = n - n1;
    PIGP_TRY(chol_rec(A, ld, n1, n2 + m_below, invd, info, base, st));
    {
        // A22 (lower) and every row below it:  C -= A21 * A21^T
        GemmDesc g{};
        g.M = (int)(n2 + m_below); g.N = (int)n2; g.K = (int)n1;
        g.alpha = -1.0; g.beta = 1.0;
        g.A = A + n1 * ld; g.lda = ld; g.a_kcontig = 1;
        g.B = A + n1 * ld; g.ldb = ld; g.b_kcontig = 1;
        g.C = A + n1 * ld + n1; g.ldc = ld;
        g.lower_only = 1;
        PIGP_TRY(launch_gemm(g, st));
    }
    return chol_rec(A + n1 * ld + n1, ld, n2, m_below, invd + (n1 / TILE) * TILE * TILE, info, base + (int)n1, st);
}

int potrf_lower(double* A, int64_t ld, int64_t n, int64_t m_extra, double* invd, int32_t* info, cudaStream_t st) {
    if (n <= 0 || n % TILE || m_extra % TILE || m_extra < 0 || ld % 2) {
        set_error("pigp potrf: n and m_extra must be multiples of 128 and ld even");
        return PIGP_EINVAL;
    }
    return chol_rec(A, ld, n, m_extra, invd, info, 0, st);
}

// ----------------------------------------------------------------------------------------------- TRTRI + LAUUM
__global__ void k_place_diag(double* W, int64_t ld, const double* invd) {
    // diagonal tile b of W <- invd[b] (full tile, zeros above the diagonal)
    const int b = blockIdx.x;
    double* dst = W + (int64_t)b * TILE * ld + (int64_t)b * TILE;
    const double* src = invd + (int64_t)b * TILE * TILE;
    for (int e = threadIdx.x; e < TILE * TILE; e += blockDim.x) dst[(int64_t)(e >> 7) * ld + (e & 127)] = src[e];
}

static int trtri_rec(const double* L, double* W, int64_t ld, int64_t n, cudaStream_t st) {
    if (n == TILE) return PIGP_OK;  // diagonal tiles were placed up front
    const int64_t n1 = (n / TILE / 2) * TILE, n2 = n - n1;
    PIGP_TRY(trtri_rec(L, W, ld, n1, st));
    PIGP_TRY(trtri_rec(L + n1 * ld + n1, W + n1 * ld + n1, ld, n2, st));
    double* Tt = W + n1;  // scratch in the strictly upper part of W: Tt (n1 x n2) = W11^T * L21^T
    {
        GemmDesc g{};
        g.M = (int)n1; g.N = (int)n2; g.K = (int)n1;
        g.alpha = 1.0; g.beta = 0.0;
        g.A = W; g.lda = ld; g.a_kcontig = 0;  // A(m,k) = W11[k][m], non-zero for k >= m
        g.B = L + n1 * ld; g.ldb = ld; g.b_kcontig = 1;  // B(n,k) = L21[n][k]
        g.C = Tt; g.ldc = ld;
        g.kmode = 1;
        PIGP_TRY(launch_gemm(g, st));
    }
    {
        // W21 = -W22 * T,  T[k][n] = Tt[n][k]
        GemmDesc g{};
        g.M = (int)n2; g.N = (int)n1; g.K = (int)n2;
        g.alpha = -1.0; g.beta = 0.0;
        g.A = W + n1 * ld + n1; g.lda = ld; g.a_kcontig = 1;  // A(m,k) = W22[m][k], non-zero for k <= m
        g.B = Tt; g.ldb = ld; g.b_kcontig = 1;
        g.C = W + n1 * ld; g.ldc = ld;
        g.kmode = 2;
        PIGP_TRY(launch_gemm(g, st));
    }
    return PIGP_OK;
}

int potri_lower(const double* L, int64_t ld, int64_t n, const double* invd, double* W, double* X, cudaStream_t st) {
    if (n <= 0 || n % TILE || ld % 2) {
        set_error("pigp potri: n must be a multiple of 128 and ld even");
        return PIGP_EINVAL;
    }
    {
        ProfScope prof(PROF_MISC, st);
        k_place_diag<<<(unsigned)(n / TILE), 256, 0, st>>>(W, ld, invd);
        count_launch();
    }
    PIGP_CUDA(cudaGetLastError());
    PIGP_TRY(trtri_rec(L, W, ld, n, st));
    // X (lower) = W^T W :  X_ij = sum_{k >= i} W[k][i] W[k][j]
    GemmDesc g{};
    g.M = (int)n; g.N = (int)n; g.K = (int)n;
    g.alpha = 1.0; g.beta = 0.0;
    g.A = W; g.lda = ld; g.a_kcontig = 0;
    g.B = W; g.ldb = ld; g.b_kcontig = 0;
    g.C = X; g.ldc = ld;
    g.lower_only = 1;
    g.kmode = 1;
    return launch_gemm(g, st);
}

// ----------------------------------------------------------------------------------------------- reductions / BLAS-2
// out2[0] = sum_{i<n} log A[i][i],  out2[1] = sum_{j<n} v[j]^2   (single CTA, fixed order: deterministic)
__global__ void __launch_bounds__(1024) k_logdet_quad(const double* A, int64_t ld, int64_t n, const double* v, double* out2) {
    __shared__ double sh[2][32];
    double s0 = 0.0, s1 = 0.0;
    for (int64_t i = threadIdx.x; i < n; i += blockDim.x) {
        s0 += log(A[i * ld + i]);
        const double x = v[i];
        s1 = fma(x, x, s1);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        s0 += __shfl_xor_sync(0xffffffffu, s0, o);
        s1 += __shfl_xor_sync(0xffffffffu, s1, o);
    }
    if ((threadIdx.x & 31) == 0) { sh[0][threadIdx.x >> 5] = s0; sh[1][threadIdx.x >> 5] = s1; }
    __syncthreads();
    if (threadIdx.x < 32) {
        s0 = sh[0][threadIdx.x];
        s1 = sh[1][threadIdx.x];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            s0 += __shfl_xor_sync(0xffffffffu, s0, o);
            s1 += __shfl_xor_sync(0xffffffffu, s1, o);
        }
        if (threadIdx.x == 0) { out2[0] = s0; out2[1] = s1; }
    }
}

int launch_logdet_quad(const double* A, int64_t ld, int64_t n, const double* v, double* out2, cudaStream_t st) {
    ProfScope prof(PROF_MISC, st);
    k_logdet_quad<<<1, 1024, 0, st>>>(A, ld, n, v, out2);
    count_launch();
    PIGP_CUDA(cudaGetLastError());
    return PIGP_OK;
}

// y[j] = sum_{i >= j} W[i][j] x[i].  Pass 1: CTA (column block, row chunk) sums its rows into part[chunk][j]
// (coalesced 1 KB row segments, 4 rows in flight per thread); pass 2 adds the chunks in order (deterministic).
constexpr int TRMV_CHUNK = 1024;
__global__ void __launch_bounds__(256) k_trmv_lower_t(const double* W, int64_t ld, int64_t n, const double* x, double* part) {
    __shared__ double sh[2][128];
    const int c = threadIdx.x & 127, h = threadIdx.x >> 7;
    const int64_t j0 = (int64_t)blockIdx.x * 128, j = j0 + c;
    const int64_t r0 = max((long long)blockIdx.y * TRMV_CHUNK, (long long)j0);
    const int64_t r1 = min(((long long)blockIdx.y + 1) * TRMV_CHUNK, (long long)n);
    double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;
    if (j < n) {
        int64_t i = r0 + h;
        for (; i + 6 < r1; i += 8) {
            const double w0 = W[i * ld + j], w1 = W[(i + 2) * ld + j], w2 = W[(i + 4) * ld + j], w3 = W[(i + 6) * ld + j];
            if (i >= j) s0 = fma(w0, x[i], s0);
            if (i + 2 >= j) s1 = fma(w1, x[i + 2], s1);
            if (i + 4 >= j) s2 = fma(w2, x[i + 4], s2);
            if (i + 6 >= j) s3 = fma(w3, x[i + 6], s3);
        }
        for (; i < r1; i += 2)
            if (i >= j) s0 = fma(W[i * ld + j], x[i], s0);
    }
    sh[h][c] = (s0 + s1) + (s2 + s3);
    __syncthreads();
    if (h == 0 && j < n) part[(int64_t)blockIdx.y * n + j] = sh[0][c] + sh[1][c];
}
__global__ void __launch_bounds__(256) k_sum_chunks(const double* part, int64_t n, int n_chunks, double* y) {
    const int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n) return;
    double s = 0.0;
    for (int c = (int)(j / TRMV_CHUNK); c < n_chunks; ++c) s += part[(int64_t)c * n + j];
    y[j] = s;
}

int launch_trmv_lower_t(const double* W, int64_t ld, int64_t n, const double* x, double* y, double* part, cudaStream_t st) {
    const int n_chunks = (int)((n + TRMV_CHUNK - 1) / TRMV_CHUNK);
    ProfScope prof(PROF_MISC, st);
    k_trmv_lower_t<<<dim3((unsigned)((n + 127) / 128), (unsigned)n_chunks), 256, 0, st>>>(W, ld, n, x, part);
    k_sum_chunks<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(part, n, n_chunks, y);
    count_launch(2);
    PIGP_CUDA(cudaGetLastError());
    return PIGP_OK;
}

// y[i] = sum_j A[i][j] x[j]: one warp per row
__global__ void __launch_bounds__(256) k_gemv(const double* A, int64_t ld, int64_t m, int64_t n, const double* x, double* y) {
    const int64_t row = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5);
    if (row >= m) return;
    const int lane = threadIdx.x & 31;
    double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;
    const double* a = A + row * ld;
    int64_t j = lane;
    for (; j + 96 < n; j += 128) {
        s0 = fma(a[j], x[j], s0);
        s1 = fma(a[j + 32], x[j + 32], s1);
        s2 = fma(a[j + 64], x[j + 64], s2);
        s3 = fma(a[j + 96], x[j + 96], s3);
    }
    for (; j < n; j += 32) s0 = fma(a[j], x[j], s0);
    double s = (s0 + s1) + (s2 + s3);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (lane == 0) y[row] = s;
}

int launch_gemv(const double* A, int64_t ld, int64_t m, int64_t n, const double* x, double* y, cudaStream_t st) {
    if (m <= 0) return PIGP_OK;
    ProfScope prof(PROF_MISC, st);
    k_gemv<<<(unsigned)((m + 7) / 8), 256, 0, st>>>(A, ld, m, n, x, y);
    count_launch();
    PIGP_CUDA(cudaGetLastError());
    return PIGP_OK;
}

}  // namespace pigp
