// extern "C" surface of libpigp.so: plans, solver workspace, NLL / gradient / posterior drivers.
// See include/pigp.h for the contract and the reference interfaces each entry point replaces.
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>

#include "pigp_internal.cuh"

namespace pigp {

static thread_local std::string t_error;
void set_error(const std::string& msg) { t_error = msg; }
std::atomic<long long> g_launches{0};

bool g_prof_on = false;
struct ProfRec { int cls; cudaEvent_t e0, e1; double flops; int m, n, k, mode; cudaStream_t st; };
static int g_note[4] = {0, 0, 0, 0};
void prof_note(int m, int n, int k, int mode) { g_note[0] = m; g_note[1] = n; g_note[2] = k; g_note[3] = mode; }
static std::vector<ProfRec> g_prof;
static std::mutex g_prof_mu;
void prof_push(int cls, cudaStream_t st, bool begin, double flops) {
    std::lock_guard<std::mutex> lk(g_prof_mu);
    if (begin) {
        ProfRec r{cls, nullptr, nullptr, flops, g_note[0], g_note[1], g_note[2], g_note[3], st};
        g_note[0] = g_note[1] = g_note[2] = g_note[3] = 0;
        cudaEventCreate(&r.e0);
        cudaEventCreate(&r.e1);
        cudaEventRecord(r.e0, st);
        g_prof.push_back(r);
    } else {
        for (size_t i = g_prof.size(); i-- > 0;)
            if (g_prof[i].cls == cls && g_prof[i].st == st && g_prof[i].e1 != nullptr) { cudaEventRecord(g_prof[i].e1, st); break; }
    }
}

static void add_tiles(std::vector<AsmTile>& out, int64_t r0, int64_t r1, int64_t c0, int64_t c1, int desc, int flags) {
    for (int64_t r = r0; r < r1; r += ASM_TR) {
        const int nr = (int)std::min<int64_t>(ASM_TR, r1 - r);
        for (int64_t c = c0; c < c1; c += ASM_TC) {
            if ((flags & ASM_LOWER) && r + nr - 1 < c) continue;  // entirely above the diagonal
            const int nc = (int)std::min<int64_t>(ASM_TC, c1 - c);
            out.push_back(AsmTile{(int32_t)r, (int32_t)c, nr, nc, desc, flags});
        }
    }
}

void build_lower_tiles_owned(const pigp_plan* p, int rank, int world, std::vector<AsmTile>& out) {
    const int nb = p->n_row_blocks;
    for (int i = 0; i < nb; ++i)
        for (int j = 0; j <= i; ++j) {
            const pigp_block_desc& b = p->table[(size_t)j * nb + i];  // lower block (i, j) = transpose of table[j][i]
            const int desc = b.n_terms > 0 ? j * nb + i : -1;
            const int flags = (i == j) ? ASM_LOWER : ASM_SWAP;
            const int64_t r1 = p->sec_row[i + 1], c0 = p->sec_row[j], c1 = p->sec_row[j + 1];
            for (int64_t r = p->sec_row[i]; r < r1;) {
                const int nr = (int)std::min<int64_t>(std::min<int64_t>(ASM_TR, r1 - r), TILE - r % TILE);
                if ((int)((r / TILE) % world) == rank)
                    for (int64_t c = c0; c < c1; c += ASM_TC) {
                        if ((flags & ASM_LOWER) && r + nr - 1 < c) continue;
                        out.push_back(AsmTile{(int32_t)r, (int32_t)c, nr, (int)std::min<int64_t>(ASM_TC, c1 - c), desc, flags});
                    }
                r += nr;
            }
        }
}

static int upload_tiles(const std::vector<AsmTile>& v, AsmTile** dev, int64_t* n) {
    *n = (int64_t)v.size();
    *dev = nullptr;
    if (v.empty()) return PIGP_OK;
    PIGP_CUDA(cudaMalloc(dev, v.size() * sizeof(AsmTile)));
    PIGP_CUDA(cudaMemcpy(*dev, v.data(), v.size() * sizeof(AsmTile), cudaMemcpyHostToDevice));
    return PIGP_OK;
}

// host [n][dim] (array of points) -> pinned [dim][n] -> device
static int upload_points(pigp_plan* p, const double* host, int64_t n, double* dev, cudaStream_t st) {
    const int64_t bytes = n * p->dim * (int64_t)sizeof(double);
    if (bytes > p->h_pin_bytes) {
        if (p->h_pin) cudaFreeHost(p->h_pin);
        p->h_pin = nullptr;
        PIGP_CUDA(cudaMallocHost(&p->h_pin, bytes));
        p->h_pin_bytes = bytes;
    }
    for (int64_t i = 0; i < n; ++i)
        for (int d = 0; d < p->dim; ++d) p->h_pin[d * n + i] = host[i * p->dim + d];
    PIGP_CUDA(cudaMemcpyAsync(dev, p->h_pin, bytes, cudaMemcpyHostToDevice, st));
    PIGP_CUDA(cudaStreamSynchronize(st));  // the pinned buffer is reused by the next call
    return PIGP_OK;
}

__global__ void k_set_yrow(double* A, int64_t ld, int64_t n, int64_t npad, int64_t yrow, int block_rows, const double* y,
                           double huge_diag) {
    // rows [yrow, yrow + block_rows): first row = [y, 0 ..], others zero.  When the row sits inside the padded
    // square (yrow < npad) only that single row is written and its diagonal becomes huge_diag.
    const int64_t total = (int64_t)block_rows * npad;
    for (int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
        const int64_t r = e / npad, c = e % npad;
        double v = 0.0;
        if (r == 0 && c < n) v = y[c];
        if (yrow < npad) {
            if (c > yrow) continue;  // lower storage
            if (c == yrow) v = huge_diag;
        }
        A[(yrow + r) * ld + c] = v;
    }
}

__global__ void k_rowsumsq_sub(const double* Vt, int64_t ld, int64_t m, int64_t n, const double* kdiag, double* var) {
    const int64_t row = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5);
    if (row >= m) return;
    const int lane = threadIdx.x & 31;
    double s = 0.0;
    for (int64_t j = lane; j < n; j += 32) {
        const double x = Vt[row * ld + j];
        s = fma(x, x, s);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (lane == 0) var[row] = kdiag[row] - s;
}

__global__ void k_take_diag(const double* T, int64_t ld, int64_t m, double* out) {
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i < m) out[i] = T[i * ld + i];
}

}  // namespace pigp

using namespace pigp;

struct pigp_solver {
    pigp_plan* plan = nullptr;
    pigp_dsolver* ds = nullptr;  // world = 1 instance of the block-cyclic solver: NLL and gradient run there
    int64_t n = 0, npad = 0;
    int64_t yrow = 0;        // row of the factorisation buffer that carries y
    int64_t mrow0 = 0;       // first row available to the mixed (test x train) block
    double* A = nullptr;     // posterior only (allocated on first use): (a_rows x npad) row-major, K -> L, y row, mixed rows
    int64_t a_rows = 0;
    double* invd = nullptr;  // npad/128 inverse diagonal tiles
    double* out2 = nullptr;  // logdet, quad
    int32_t* info = nullptr;
    double* T = nullptr;     // test covariance scratch
    int64_t t_elems = 0;
    // staging for the _host entry points
    double* d_theta = nullptr;
    double* d_y = nullptr;
    double* d_res = nullptr;  // nll + grad
    double* h_res = nullptr;  // pinned
    double* d_mu = nullptr;
    int64_t mu_elems = 0;
};

static cudaStream_t as_stream(void* s) { return reinterpret_cast<cudaStream_t>(s); }

extern "C" {

int pigp_abi_version(void) { return PIGP_ABI_VERSION; }
const char* pigp_last_error(void) { return t_error.c_str(); }
int64_t pigp_launch_count(void) { return (int64_t)g_launches.load(); }

int pigp_profile_start(void) {
    std::lock_guard<std::mutex> lk(g_prof_mu);
    g_prof.clear();
    g_prof_on = true;
    return PIGP_OK;
}

int pigp_profile_stop(double* ms_out, int64_t* launches_out, double* flops_out) {
    g_prof_on = false;
    PIGP_CUDA(cudaDeviceSynchronize());
    std::lock_guard<std::mutex> lk(g_prof_mu);
    for (int c = 0; c < PIGP_PROF_CLASSES; ++c) {
        if (ms_out) ms_out[c] = 0.0;
        if (launches_out) launches_out[c] = 0;
        if (flops_out) flops_out[c] = 0.0;
    }
    FILE* dump = nullptr;
    if (const char* path = getenv("PIGP_PROF_DUMP")) dump = fopen(path, "w");
    if (dump) fprintf(dump, "class,ms,flops,m,n,k,mode,start_ms,stream\n");
    for (auto& r : g_prof) {
        float ms = 0.f, t0 = 0.f;
        cudaEventElapsedTime(&ms, r.e0, r.e1);
        if (dump) {
            cudaEventElapsedTime(&t0, g_prof.front().e0, r.e0);
            fprintf(dump, "%d,%.6f,%.0f,%d,%d,%d,%d,%.6f,%p\n", r.cls, ms, r.flops, r.m, r.n, r.k, r.mode, t0, (void*)r.st);
        }
        if (ms_out) ms_out[r.cls] += ms;
        if (launches_out) launches_out[r.cls] += 1;
        if (flops_out) flops_out[r.cls] += r.flops;
    }
    for (auto& r : g_prof) {
        cudaEventDestroy(r.e0);
        cudaEventDestroy(r.e1);
    }
    if (dump) fclose(dump);
    g_prof.clear();
    return PIGP_OK;
}

int pigp_set_device(int device) {
    PIGP_CUDA(cudaSetDevice(device));
    return PIGP_OK;
}

void pigp_plan_destroy(pigp_plan* p) {
    if (!p) return;
    if (p->d_pts_col && p->d_pts_col != p->d_pts_row) cudaFree(p->d_pts_col);
    cudaFree(p->d_pts_row);
    cudaFree(p->d_table);
    cudaFree(p->d_tiles_full);
    cudaFree(p->d_tiles_lower);
    cudaFree(p->d_theta_stage);
    if (p->h_pin) cudaFreeHost(p->h_pin);
    delete p;
}

int pigp_plan_create(const pigp_plan_desc* d, pigp_plan** out) {
    if (!d || !out) { set_error("pigp_plan_create: null argument"); return PIGP_EINVAL; }
    *out = nullptr;
    if (d->dim < 1 || d->dim > 3 || d->n_groups < 1 || d->n_groups > PIGP_MAX_GROUPS || d->n_row_blocks < 1 ||
        (!d->symmetric && d->n_col_blocks < 1) || !d->sec_row || !d->pts_row_host || !d->table) {
        set_error("pigp_plan_create: bad dimensions or null arrays");
        return PIGP_EINVAL;
    }
    pigp_plan* p = new pigp_plan();
    p->dim = d->dim;
    p->product_form = d->product_form ? 1 : 0;
    p->n_groups = d->n_groups;
    p->symmetric = d->symmetric ? 1 : 0;
    p->n_row_blocks = d->n_row_blocks;
    p->n_col_blocks = p->symmetric ? d->n_row_blocks : d->n_col_blocks;
    p->sec_row.assign(d->sec_row, d->sec_row + d->n_row_blocks + 1);
    if (p->symmetric) p->sec_col = p->sec_row;
    else p->sec_col.assign(d->sec_col, d->sec_col + d->n_col_blocks + 1);
    p->rows = p->sec_row.back();
    p->cols = p->sec_col.back();
    for (int k = 0; k < 3; ++k) p->lbox[k] = d->lbox[k];
    p->noise_lo_block = d->noise_lo_block;
    p->noise_hi_block = d->noise_hi_block;
    p->theta_len = d->n_groups * (1 + d->dim);
    auto fail = [&](const char* msg) {
        set_error(msg);
        pigp_plan_destroy(p);
        return PIGP_EINVAL;
    };
    for (size_t i = 0; i + 1 < p->sec_row.size(); ++i)
        if (p->sec_row[i + 1] < p->sec_row[i]) return fail("pigp_plan_create: sec_row must be non-decreasing");
    for (size_t i = 0; i + 1 < p->sec_col.size(); ++i)
        if (p->sec_col[i + 1] < p->sec_col[i]) return fail("pigp_plan_create: sec_col must be non-decreasing");
    if (p->rows <= 0 || p->cols <= 0 || p->rows > (1 << 30) || p->cols > (1 << 30))
        return fail("pigp_plan_create: empty or oversized matrix");
    if (p->noise_lo_block >= 0) {
        if (!p->symmetric || p->noise_hi_block < p->noise_lo_block || p->noise_hi_block >= p->n_row_blocks)
            return fail("pigp_plan_create: bad noise block range");
        p->noise_lo = p->sec_row[p->noise_lo_block];
        p->noise_hi = p->sec_row[p->noise_hi_block + 1];
        p->theta_len += 1;
    }
    const int nrb = p->n_row_blocks, ncb = p->n_col_blocks;
    p->table.assign(d->table, d->table + (size_t)nrb * ncb);
    for (int i = 0; i < nrb; ++i)
        for (int j = (p->symmetric ? i : 0); j < ncb; ++j) {
            const pigp_block_desc& b = p->table[(size_t)i * ncb + j];
            if (b.n_terms < 0 || b.n_terms > PIGP_MAX_TERMS) return fail("pigp_plan_create: n_terms out of range");
            for (int t = 0; t < b.n_terms; ++t) {
                if (b.terms[t].group < 0 || b.terms[t].group >= p->n_groups) return fail("pigp_plan_create: term group out of range");
                if (t && b.terms[t].group < b.terms[t - 1].group) return fail("pigp_plan_create: terms must be sorted by group");
                int active = 0, total = 0;
                for (int k = 0; k < p->dim; ++k) {
                    if (b.terms[t].order[k] < -1 || b.terms[t].order[k] > 4) return fail("pigp_plan_create: derivative order out of range");
                    if (b.terms[t].order[k] >= 0) { ++active; total += b.terms[t].order[k]; }
                }
                if (total > 4) return fail("pigp_plan_create: total derivative order of a term exceeds 4");
                if (!p->product_form && p->dim > 1 && active != 1)
                    return fail("pigp_plan_create: additive kernel form: every term must act on exactly one dimension (order -1 elsewhere)");
            }
        }
    if (cudaGetDevice(&p->device) != cudaSuccess) return fail("pigp_plan_create: no CUDA device");

    int rc = PIGP_OK;
    auto cuda_ok = [&](cudaError_t e, const char* what) {
        if (e != cudaSuccess && rc == PIGP_OK) {
            set_error(std::string(what) + ": " + cudaGetErrorString(e));
            rc = PIGP_ECUDA;
        }
    };
    cuda_ok(cudaMalloc(&p->d_pts_row, sizeof(double) * p->dim * p->rows), "cudaMalloc pts_row");
    if (p->symmetric) p->d_pts_col = p->d_pts_row;
    else cuda_ok(cudaMalloc(&p->d_pts_col, sizeof(double) * p->dim * p->cols), "cudaMalloc pts_col");
    cuda_ok(cudaMalloc(&p->d_table, sizeof(pigp_block_desc) * p->table.size()), "cudaMalloc table");
    cuda_ok(cudaMalloc(&p->d_theta_stage, sizeof(double) * MAX_THETA), "cudaMalloc theta");
    if (rc == PIGP_OK)
        cuda_ok(cudaMemcpy(p->d_table, p->table.data(), sizeof(pigp_block_desc) * p->table.size(), cudaMemcpyHostToDevice), "copy table");
    if (rc == PIGP_OK) rc = upload_points(p, d->pts_row_host, p->rows, p->d_pts_row, 0);
    if (rc == PIGP_OK && !p->symmetric) {
        if (!d->pts_col_host || !d->sec_col) { pigp_plan_destroy(p); set_error("pigp_plan_create: rectangular plan needs column points"); return PIGP_EINVAL; }
        rc = upload_points(p, d->pts_col_host, p->cols, p->d_pts_col, 0);
    }
    if (rc == PIGP_OK) {
        std::vector<AsmTile> full, lower;
        auto desc_of = [&](int i, int j) { return p->table[(size_t)i * ncb + j].n_terms > 0 ? i * ncb + j : -1; };
        for (int i = 0; i < nrb; ++i)
            for (int j = 0; j < ncb; ++j) {
                const int64_t r0 = p->sec_row[i], r1 = p->sec_row[i + 1], c0 = p->sec_col[j], c1 = p->sec_col[j + 1];
                if (!p->symmetric) add_tiles(full, r0, r1, c0, c1, desc_of(i, j), 0);
                if (p->symmetric && i == j) add_tiles(lower, r0, r1, c0, c1, desc_of(i, i), ASM_LOWER);
                if (p->symmetric && i > j) add_tiles(lower, r0, r1, c0, c1, desc_of(j, i), ASM_SWAP);  // lower block = transpose of table[j][i]
            }
        if (p->symmetric) {
            // full layout = the lower tiles, each also stored transposed: half the evaluations, exact symmetry
            full = lower;
            for (AsmTile& t : full) t.flags |= ASM_MIRROR;
        }
        rc = upload_tiles(full, &p->d_tiles_full, &p->n_tiles_full);
        if (rc == PIGP_OK) rc = upload_tiles(lower, &p->d_tiles_lower, &p->n_tiles_lower);
    }
    if (rc != PIGP_OK) { pigp_plan_destroy(p); return rc; }
    *out = p;
    return PIGP_OK;
}

int64_t pigp_plan_rows(const pigp_plan* p) { return p ? p->rows : 0; }
int64_t pigp_plan_cols(const pigp_plan* p) { return p ? p->cols : 0; }
int32_t pigp_plan_theta_len(const pigp_plan* p) { return p ? p->theta_len : 0; }

int pigp_plan_set_points_host(pigp_plan* p, int side, const double* pts_host, void* stream) {
    if (!p || !pts_host) { set_error("pigp_plan_set_points_host: null argument"); return PIGP_EINVAL; }
    if (side == 0 || p->symmetric) return upload_points(p, pts_host, p->rows, p->d_pts_row, as_stream(stream));
    return upload_points(p, pts_host, p->cols, p->d_pts_col, as_stream(stream));
}

int pigp_assemble(const pigp_plan* p, const double* theta_dev, double eps, int add_diag, double* K_dev, int64_t ld,
                  int layout, void* stream) {
    if (!p || !theta_dev || !K_dev || ld < p->cols) { set_error("pigp_assemble: bad argument"); return PIGP_EINVAL; }
    if (add_diag && !p->symmetric) { set_error("pigp_assemble: add_diag needs a symmetric plan"); return PIGP_EINVAL; }
    if (layout == PIGP_LAYOUT_LOWER) {
        if (!p->symmetric) { set_error("pigp_assemble: LOWER layout needs a symmetric plan"); return PIGP_EINVAL; }
        return launch_assemble(p, p->d_tiles_lower, p->n_tiles_lower, theta_dev, eps, add_diag, K_dev, ld, as_stream(stream));
    }
    return launch_assemble(p, p->d_tiles_full, p->n_tiles_full, theta_dev, eps, add_diag, K_dev, ld, as_stream(stream));
}

int pigp_assemble_host(pigp_plan* p, const double* theta_host, double eps, int add_diag, double* K_host, int layout) {
    if (!p || !theta_host || !K_host) { set_error("pigp_assemble_host: null argument"); return PIGP_EINVAL; }
    double* K = nullptr;
    const size_t bytes = sizeof(double) * (size_t)p->rows * p->cols;
    PIGP_CUDA(cudaMalloc(&K, bytes));
    int rc = PIGP_OK;
    if (cudaMemcpy(p->d_theta_stage, theta_host, sizeof(double) * p->theta_len, cudaMemcpyHostToDevice) != cudaSuccess) rc = PIGP_ECUDA;
    if (rc == PIGP_OK && layout == PIGP_LAYOUT_LOWER && cudaMemset(K, 0, bytes) != cudaSuccess) rc = PIGP_ECUDA;
    if (rc == PIGP_OK) rc = pigp_assemble(p, p->d_theta_stage, eps, add_diag, K, p->cols, layout, nullptr);
    if (rc == PIGP_OK && cudaMemcpy(K_host, K, bytes, cudaMemcpyDeviceToHost) != cudaSuccess) {
        set_error("pigp_assemble_host: device to host copy failed");
        rc = PIGP_ECUDA;
    }
    cudaFree(K);
    return rc;
}

// ------------------------------------------------------------------------------------------------ solver
void pigp_solver_destroy(pigp_solver* s) {
    if (!s) return;
    pigp_dsolver_destroy(s->ds);
    cudaFree(s->A); cudaFree(s->invd); cudaFree(s->out2); cudaFree(s->info); cudaFree(s->T); cudaFree(s->d_theta); cudaFree(s->d_y);
    cudaFree(s->d_res); cudaFree(s->d_mu);
    if (s->h_res) cudaFreeHost(s->h_res);
    delete s;
}

static int ensure_rows(pigp_solver* s, int64_t rows) {
    if (rows <= s->a_rows) return PIGP_OK;
    if (s->A) cudaFree(s->A);
    s->A = nullptr;
    s->a_rows = 0;
    PIGP_CUDA(cudaMalloc(&s->A, sizeof(double) * (size_t)rows * s->npad));
    s->a_rows = rows;
    return PIGP_OK;
}

int pigp_solver_create(pigp_plan* plan, pigp_solver** out) {
    if (!plan || !out || !plan->symmetric) { set_error("pigp_solver_create: needs a symmetric training plan"); return PIGP_EINVAL; }
    *out = nullptr;
    pigp_solver* s = new pigp_solver();
    s->plan = plan;
    s->n = plan->rows;
    s->npad = round_up(s->n, TILE);
    if (s->n < s->npad) { s->yrow = s->n; s->mrow0 = s->npad; }          // y rides in the first padding row
    else { s->yrow = s->npad; s->mrow0 = s->npad + TILE; }                // no padding row: y gets its own block
    int rc = pigp_dsolver_create(plan, 0, 1, &s->ds);  // A (posterior only) is allocated on first use
    auto cuda_ok = [&](cudaError_t e, const char* what) {
        if (e != cudaSuccess && rc == PIGP_OK) { set_error(std::string(what) + ": " + cudaGetErrorString(e)); rc = PIGP_ECUDA; }
    };
    const int64_t nt = s->npad / TILE;
    cuda_ok(cudaMalloc(&s->invd, sizeof(double) * nt * TILE * TILE), "cudaMalloc invd");
    cuda_ok(cudaMalloc(&s->out2, sizeof(double) * 2), "cudaMalloc out2");
    cuda_ok(cudaMalloc(&s->info, sizeof(int32_t)), "cudaMalloc info");
    cuda_ok(cudaMalloc(&s->d_theta, sizeof(double) * MAX_THETA), "cudaMalloc theta");
    cuda_ok(cudaMalloc(&s->d_y, sizeof(double) * s->n), "cudaMalloc y");
    cuda_ok(cudaMalloc(&s->d_res, sizeof(double) * (1 + MAX_THETA)), "cudaMalloc res");
    cuda_ok(cudaMallocHost(&s->h_res, sizeof(double) * (2 + 2 * MAX_THETA) + sizeof(double) * s->n), "cudaMallocHost res");
    if (rc != PIGP_OK) { pigp_solver_destroy(s); return rc; }
    *out = s;
    return PIGP_OK;
}

// assemble K (lower, jitter added), pad, place y, factor with `extra_rows` more rows under the square
static int factor(pigp_solver* s, const double* theta, const double* y, double eps, int64_t extra_rows, int32_t* info,
                  cudaStream_t st) {
    const pigp_plan* p = s->plan;
    const int64_t ld = s->npad;
    PIGP_CUDA(cudaMemsetAsync(info, 0, sizeof(int32_t), st));
    PIGP_TRY(launch_assemble(p, p->d_tiles_lower, p->n_tiles_lower, theta, eps, 1, s->A, ld, st));
    PIGP_TRY(launch_pad(s->A, ld, s->n, s->n, s->npad, s->npad, 1, 1, st));
    const int yblock = (s->yrow < s->npad) ? 1 : TILE;
    {
        ProfScope prof(PROF_MISC, st);
        k_set_yrow<<<(unsigned)std::min<int64_t>((yblock * s->npad + 255) / 256, 1184), 256, 0, st>>>(
            s->A, ld, s->n, s->npad, s->yrow, yblock, y, 1e300);
        count_launch();
    }
    PIGP_CUDA(cudaGetLastError());
    return potrf_lower(s->A, ld, s->npad, (s->mrow0 - s->npad) + extra_rows, s->invd, info, st);
}

int pigp_nll(pigp_solver* s, const double* theta_dev, const double* y_dev, double eps, double* out_dev, int32_t* info_dev,
             void* stream) {
    if (!s || !theta_dev || !y_dev || !out_dev) { set_error("pigp_nll: null argument"); return PIGP_EINVAL; }
    return pigp_dsolver_nll_grad(s->ds, theta_dev, y_dev, eps, out_dev, nullptr, info_dev ? info_dev : s->info, stream);
}

int pigp_nll_grad(pigp_solver* s, const double* theta_dev, const double* y_dev, double eps, double* nll_dev, double* grad_dev,
                  int32_t* info_dev, void* stream) {
    if (!s || !theta_dev || !y_dev || !nll_dev || !grad_dev) { set_error("pigp_nll_grad: null argument"); return PIGP_EINVAL; }
    return pigp_dsolver_nll_grad(s->ds, theta_dev, y_dev, eps, nll_dev, grad_dev, info_dev ? info_dev : s->info, stream);
}

int pigp_nll_grad_host(pigp_solver* s, const double* theta_host, const double* pts_host, const double* y_host, double eps,
                       int want_grad, double* nll_host, double* grad_host, int32_t* info_host) {
    if (!s || !theta_host || !y_host || !nll_host || (want_grad && !grad_host)) { set_error("pigp_nll_grad_host: null argument"); return PIGP_EINVAL; }
    pigp_plan* p = s->plan;
    const int P = p->theta_len;
    cudaStream_t st = 0;
    if (pts_host) PIGP_TRY(upload_points(p, pts_host, p->rows, p->d_pts_row, st));
    double* h_theta = s->h_res + 2 + MAX_THETA;
    double* h_y = h_theta + MAX_THETA;
    std::memcpy(h_theta, theta_host, sizeof(double) * P);
    std::memcpy(h_y, y_host, sizeof(double) * s->n);
    PIGP_CUDA(cudaMemcpyAsync(s->d_theta, h_theta, sizeof(double) * P, cudaMemcpyHostToDevice, st));
    PIGP_CUDA(cudaMemcpyAsync(s->d_y, h_y, sizeof(double) * s->n, cudaMemcpyHostToDevice, st));
    if (want_grad) PIGP_TRY(pigp_nll_grad(s, s->d_theta, s->d_y, eps, s->d_res, s->d_res + 1, s->info, st));
    else PIGP_TRY(pigp_nll(s, s->d_theta, s->d_y, eps, s->d_res, s->info, st));
    PIGP_CUDA(cudaMemcpyAsync(s->h_res, s->d_res, sizeof(double) * (1 + (want_grad ? P : 0)), cudaMemcpyDeviceToHost, st));
    int32_t* h_info = reinterpret_cast<int32_t*>(s->h_res + 1 + MAX_THETA);
    PIGP_CUDA(cudaMemcpyAsync(h_info, s->info, sizeof(int32_t), cudaMemcpyDeviceToHost, st));
    PIGP_CUDA(cudaStreamSynchronize(st));
    *nll_host = s->h_res[0];
    if (want_grad) std::memcpy(grad_host, s->h_res + 1, sizeof(double) * P);
    if (info_host) *info_host = *h_info;
    return PIGP_OK;
}

int pigp_predict(pigp_solver* s, const pigp_plan* mixed, const pigp_plan* test, const double* theta_dev, const double* y_dev,
                 double eps, double* mu_dev, double* cov_dev, int want_full_cov, int32_t* info_dev, void* stream) {
    if (!s || !mixed || !test || !theta_dev || !y_dev || !mu_dev || !cov_dev) { set_error("pigp_predict: null argument"); return PIGP_EINVAL; }
    if (mixed->symmetric || mixed->cols != s->n || !test->symmetric || test->rows != mixed->rows) {
        set_error("pigp_predict: mixed must be (test x train) and test symmetric over the same test points");
        return PIGP_EINVAL;
    }
    cudaStream_t st = as_stream(stream);
    int32_t* info = info_dev ? info_dev : s->info;
    const int64_t m = mixed->rows, mpad = round_up(m, TILE), ld = s->npad;
    PIGP_TRY(ensure_rows(s, s->mrow0 + mpad));
    double* Vt = s->A + s->mrow0 * ld;
    // mixed block under the square: rows = test points (first kernel argument), columns = training points
    PIGP_TRY(launch_assemble(mixed, mixed->d_tiles_full, mixed->n_tiles_full, theta_dev, 0.0, 0, Vt, ld, st));
    PIGP_TRY(launch_pad(Vt, ld, m, s->n, mpad, s->npad, 0, 0, st));
    PIGP_TRY(factor(s, theta_dev, y_dev, eps, mpad, info, st));
    // mu = K_ab K_bb^-1 y = (K_ab L^-T)(L^-1 y)
    PIGP_TRY(launch_gemv(Vt, ld, m, s->n, s->A + s->yrow * ld, mu_dev, st));
    // K_aa - V^T V
    if (s->t_elems < mpad * mpad) {
        if (s->T) cudaFree(s->T);
        s->T = nullptr; s->t_elems = 0;
        PIGP_CUDA(cudaMalloc(&s->T, sizeof(double) * (size_t)mpad * mpad));
        s->t_elems = mpad * mpad;
    }
    PIGP_TRY(launch_assemble(test, test->d_tiles_full, test->n_tiles_full, theta_dev, 0.0, 0, s->T, mpad, st));
    PIGP_TRY(launch_pad(s->T, mpad, m, m, mpad, mpad, 0, 0, st));
    if (want_full_cov) {
        GemmDesc g{};
        g.M = (int)mpad; g.N = (int)mpad; g.K = (int)s->npad;
        g.alpha = -1.0; g.beta = 1.0;
        g.A = Vt; g.lda = ld; g.a_kcontig = 1;
        g.B = Vt; g.ldb = ld; g.b_kcontig = 1;
        g.C = s->T; g.ldc = mpad;
        PIGP_TRY(launch_gemm(g, st));
        PIGP_CUDA(cudaMemcpy2DAsync(cov_dev, sizeof(double) * m, s->T, sizeof(double) * mpad, sizeof(double) * m, m,
                                    cudaMemcpyDeviceToDevice, st));
    } else {
        if (s->mu_elems < m) {
            if (s->d_mu) cudaFree(s->d_mu);
            s->d_mu = nullptr; s->mu_elems = 0;
            PIGP_CUDA(cudaMalloc(&s->d_mu, sizeof(double) * m));
            s->mu_elems = m;
        }
        k_take_diag<<<(unsigned)((m + 255) / 256), 256, 0, st>>>(s->T, mpad, m, s->d_mu);
        k_rowsumsq_sub<<<(unsigned)((m + 7) / 8), 256, 0, st>>>(Vt, ld, m, s->n, s->d_mu, cov_dev);
        count_launch(2);
        PIGP_CUDA(cudaGetLastError());
    }
    return PIGP_OK;
}

int pigp_predict_host(pigp_solver* s, pigp_plan* mixed, pigp_plan* test, const double* theta_host, const double* y_host,
                      double eps, double* mu_host, double* cov_host, int want_full_cov, int32_t* info_host) {
    if (!s || !mixed || !test || !theta_host || !y_host || !mu_host || !cov_host) { set_error("pigp_predict_host: null argument"); return PIGP_EINVAL; }
    const int P = s->plan->theta_len;
    const int64_t m = mixed->rows;
    const size_t cov_elems = want_full_cov ? (size_t)m * m : (size_t)m;
    double *d_mu = nullptr, *d_cov = nullptr;
    PIGP_CUDA(cudaMalloc(&d_mu, sizeof(double) * m));
    if (cudaMalloc(&d_cov, sizeof(double) * cov_elems) != cudaSuccess) { cudaFree(d_mu); set_error("pigp_predict_host: out of memory"); return PIGP_ENOMEM; }
    int rc = PIGP_OK;
    auto ck = [&](cudaError_t e) { if (e != cudaSuccess && rc == PIGP_OK) { set_error(cudaGetErrorString(e)); rc = PIGP_ECUDA; } };
    ck(cudaMemcpy(s->d_theta, theta_host, sizeof(double) * P, cudaMemcpyHostToDevice));
    ck(cudaMemcpy(s->d_y, y_host, sizeof(double) * s->n, cudaMemcpyHostToDevice));
    if (rc == PIGP_OK) rc = pigp_predict(s, mixed, test, s->d_theta, s->d_y, eps, d_mu, d_cov, want_full_cov, s->info, nullptr);
    if (rc == PIGP_OK) {
        ck(cudaMemcpy(mu_host, d_mu, sizeof(double) * m, cudaMemcpyDeviceToHost));
        ck(cudaMemcpy(cov_host, d_cov, sizeof(double) * cov_elems, cudaMemcpyDeviceToHost));
        int32_t inf = 0;
        ck(cudaMemcpy(&inf, s->info, sizeof(int32_t), cudaMemcpyDeviceToHost));
        if (info_host) *info_host = inf;
    }
    cudaFree(d_mu);
    cudaFree(d_cov);
    return rc;
}

// ------------------------------------------------------------------------------------------------ building blocks
int pigp_potrf_lower(double* A_dev, int64_t ld, int64_t n, int64_t m_extra, double* invd_dev, int32_t* info_dev, void* stream) {
    if (!A_dev || !invd_dev) { set_error("pigp_potrf_lower: null argument"); return PIGP_EINVAL; }
    if (info_dev) PIGP_CUDA(cudaMemsetAsync(info_dev, 0, sizeof(int32_t), as_stream(stream)));
    return potrf_lower(A_dev, ld, n, m_extra, invd_dev, info_dev, as_stream(stream));
}

int pigp_debug_potf2_stamps(long long* dev_buf) { return set_potf2_debug(dev_buf); }

int pigp_potri_lower(const double* L_dev, int64_t ld, int64_t n, const double* invd_dev, double* W_dev, double* X_dev, void* stream) {
    if (!L_dev || !invd_dev || !W_dev || !X_dev) { set_error("pigp_potri_lower: null argument"); return PIGP_EINVAL; }
    return potri_lower(L_dev, ld, n, invd_dev, W_dev, X_dev, as_stream(stream));
}

int pigp_dgemm(int M, int N, int K, double alpha, const double* A_dev, int64_t lda, int a_kcontig, const double* B_dev, int64_t ldb,
               int b_kcontig, double beta, double* C_dev, int64_t ldc, int lower_only, void* stream) {
    GemmDesc g{};
    g.M = M; g.N = N; g.K = K; g.alpha = alpha; g.beta = beta;
    g.A = A_dev; g.lda = lda; g.a_kcontig = a_kcontig;
    g.B = B_dev; g.ldb = ldb; g.b_kcontig = b_kcontig;
    g.C = C_dev; g.ldc = ldc; g.lower_only = lower_only;
    return launch_gemm(g, as_stream(stream));
}

}  // extern "C"
