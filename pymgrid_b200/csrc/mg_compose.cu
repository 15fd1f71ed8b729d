// mg_compose.cu -- composed-microgrid step (any module list) for a batch: kernels + the C-ABI of
// include/pymgrid_b200_compose.h.  Linked into libpymgrid_b200.so beside mg_engine.cu.
//
// One CTA = a tile of MGC_TILE envs, one launch = n_steps steps.  Per step:
//   1. thread e <-> env: the whole Microgrid.run dispatch (mg_compose_step.h: mgc_env_step) over the module table, state
//      rows read and written in place, reward / done / info / flags written back;
//   2. the CTA emits the tile's observation rows: a warp owns whole rows (rows w, w + 4, ... of the tile) and streams each
//      one front to back, lane l writing elements l, l + 32, ... -- every store instruction covers 256 contiguous bytes
//      and no row is written by two warps (the fused path measured whole rows from one writer ~20 % faster than split
//      rows: partial-sector merging).  An element is decoded through a per-composition table
//      (element -> module, offset; built by mgc_create) and, for time series, gathered from the pre-normalised pool,
//      whose windows are contiguous slices that stay in L2.
// This is the general path; the pymgrid25 / MicrogridGenerator module set has kernels of its own (mg_engine.cu), which is
// where the throughput work lives.  HBM-bound like them: per env-step algorithmic bytes = 8 n_act + state r/w + 9 +
// 8 obs_dim.
//
// The same file compiles for the HOST with -DMGC_HOSTSIM (g++ -x c++, tests/hostsim/): launches become loops over tiles
// and threads, pointers are host pointers.  That build is test infrastructure -- it lets the CPU suite drive this ABI,
// the layout validation and the per-env arithmetic against the reference's recorded outputs; it is never loaded by the
// package (pymgrid_b200/_cabi.py loads libpymgrid_b200.so only and fails loudly without it).
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <new>

#ifndef MGC_HOSTSIM
#include <cuda_runtime.h>
#endif

#include "mg_compose_step.h"

#define MGC_TILE 128

struct MgcLaunch {
    MgcModule mod[MGC_MAX_MODULES];
    int32_t n_mod, n_act, obs_dim, n_fstate, n_istate, cfg_stride, T, n_envs;
    int32_t n_act_modules;      // action columns of mgc_modules_step: one per non-fixed module, two per genset
    const double *cfg;
    const double *series;
    const int64_t *series_off;
    const double *series_nrm;
    int32_t *step;
    double *fstate;
    int32_t *istate;
    const int32_t *cfg_index;
    const int16_t *plist;
    int32_t n_plist, plist_width;
    const int32_t *env_initial, *env_final;
    const int32_t *elem;        // [obs_dim] module << 16 | offset inside the module's block (device copy owned by the handle)
    // the gather form of the same table (device build): an observation row is series windows plus a few state elements
    const int32_t *elem2;       // [obs_dim][2]: series element -> (module | 0x100 | forecast row << 9, offset in the block);
                                //               state element -> (0, index q into `nts`)
    const int32_t *nts;         // [n_nts] module << 16 | offset of the elements that are NOT series values (battery, genset)
    int32_t n_nts, gather;      // gather: the launch stages per-env window bases / state elements in shared memory
    int32_t hmax;               // the longest forecast horizon of the row (rows with step + hmax < T never need a fill value)
    // per call
    MgcIO io;
    int32_t mode, n_steps, ring, normalized;
};

enum { MGC_MODE_RUN = 0, MGC_MODE_RESET = 1, MGC_MODE_OBSERVE = 2, MGC_MODE_RUN_DISCRETE = 3, MGC_MODE_MODULES = 4 };

#ifdef MGC_HOSTSIM
#define MGC_DEV static inline
#else
#define MGC_DEV __device__ __forceinline__
#endif

MGC_DEV MgcView mgc_view(const MgcLaunch &P, int e) {
    MgcView V;
    V.mod = P.mod;
    V.n_mod = P.n_mod;
    V.cfg = P.cfg + (int64_t)P.cfg_index[e] * P.cfg_stride;
    V.series = P.series;
    V.series_off = P.series_off;
    V.series_nrm = P.series_nrm;
    V.T = P.T;
    return V;
}

// phase 1 of a step for env e (one thread)
MGC_DEV void mgc_owner(const MgcLaunch &P, int e, int s) {
    MgcView V = mgc_view(P, e);
    if (P.mode == MGC_MODE_RUN || P.mode == MGC_MODE_RUN_DISCRETE || P.mode == MGC_MODE_MODULES) {
        int32_t t = P.step[e];
        uint32_t flags = 0;
        double reward;
        uint8_t done;
        const int64_t slot = (int64_t)s * P.n_envs + e;
        const int act_width = (P.mode == MGC_MODE_MODULES) ? P.n_act_modules : P.n_act;
        const double *action = P.io.actions ? P.io.actions + slot * act_width : nullptr;
        double *info = P.io.info ? P.io.info + (int64_t)e * (P.n_mod * MGC_INFO_SLOTS + MGC_BALANCE_SLOTS) : nullptr;
        double *fstate = P.fstate + (int64_t)e * P.n_fstate;
        int32_t *istate = P.istate + (int64_t)e * P.n_istate;
        int normalized = P.normalized;
        double ctl[2 * MGC_MAX_MODULES];
        bool skip = false;
        if (P.mode == MGC_MODE_RUN_DISCRETE) {
            const int32_t a = P.io.dactions[P.io.dactions_const ? e : slot];
            if (a < 0 || a >= P.n_plist) {              // ValueError, envs/discrete/discrete.py:84
                flags |= MG_FLAG_BAD_ACTION;
                reward = NAN;
                done = 0;
                skip = true;
            } else if (P.T == 0 || t < P.T) {           // past the end of the series the step itself reports IndexError
                mgc_priority_control(V, t, fstate, istate, P.plist + (int64_t)a * P.plist_width * 2, P.plist_width, P.n_act,
                                     ctl, &flags);
            }
            action = ctl;
            normalized = 0;
        }
        const int final_step = P.env_final ? P.env_final[e] : (int)V.cfg[1];
        if (P.mode == MGC_MODE_MODULES) mgc_modules_step(V, t, fstate, istate, action, normalized, final_step, &reward, &done, info, &flags);
        else if (!skip) mgc_env_step(V, t, fstate, istate, action, normalized, final_step, &reward, &done, info, &flags);
        P.step[e] = t;
        P.io.reward[slot] = reward;
        P.io.done[slot] = done;
        if (P.io.flags) P.io.flags[e] = (s == 0 ? 0u : P.io.flags[e]) | flags;
    } else if (P.mode == MGC_MODE_RESET) {
        // microgrid.py:205-225: only the step moves
        if (!P.io.mask || P.io.mask[e]) P.step[e] = P.env_initial ? P.env_initial[e] : (int32_t)V.cfg[0];
    }
}

// phase 2: lane `lane` of the warp that owns row e writes its share of the row
MGC_DEV void mgc_emit_row(const MgcLaunch &P, double *obs, int e, int lane) {
    MgcView V = mgc_view(P, e);
    const int t = P.step[e];
    const double *fstate = P.fstate + (int64_t)e * P.n_fstate;
    const int32_t *istate = P.istate + (int64_t)e * P.n_istate;
    double *row = obs + (int64_t)e * P.obs_dim;
    for (int j = lane; j < P.obs_dim; j += 32) {
        const int32_t d = P.elem[j];
        row[j] = mgc_obs_element(V, d >> 16, d & 0xffff, t, fstate, istate);
    }
}

#ifndef MGC_HOSTSIM
// ---- gather emission (device build) ---------------------------------------------------------------------------------
// The decode of mgc_emit_row is the same for every row of a batch (one composition); only the window a series element
// comes from differs per env.  When the launch has `gather` set, the thread that has just stepped env e leaves in shared
// memory what the emitters need for its row: the env's step, for every series module the pool index of the window's first
// element (series_off[series_index] + t * C: two dependent loads the owner's step has just made), and the values of the
// few elements that come from module state (battery soc / charge, the four genset integers: mgc_obs_element, one thread).
// An emitter lane is then table entry -> base -> one load from the pre-normalised pool -> one streaming store, four
// elements in flight per lane, and it reads no per-env state from global memory.  Rows past the end of the series take
// mgc_obs_element (the fill row), like before.
#define MGC_STR (MGC_TILE + 1)      // row stride of the staged tables: owners write, emitters read without bank conflicts

struct MgcStage {
    int64_t *base;      // [n_mod][MGC_STR]: address of element (t, column 0) of the module's pre-normalised series
    double *val;        // [n_nts][MGC_STR]
    int32_t *t;         // [MGC_TILE]
};

__device__ __forceinline__ MgcStage mgc_stage(const MgcLaunch &P, unsigned char *smem) {
    MgcStage S;
    S.base = reinterpret_cast<int64_t *>(smem);
    S.val = reinterpret_cast<double *>(smem) + (size_t)P.n_mod * MGC_STR;
    S.t = reinterpret_cast<int32_t *>(S.val + (size_t)P.n_nts * MGC_STR);
    return S;
}
static size_t mgc_stage_bytes(const MgcLaunch &P) {
    return sizeof(double) * (size_t)(P.n_mod + P.n_nts) * MGC_STR + sizeof(int32_t) * MGC_TILE;
}

__device__ __forceinline__ void mgc_stage_env(const MgcLaunch &P, const MgcStage &S, int e, int tid) {
    const MgcView V = mgc_view(P, e);
    const int t = P.step[e];
    S.t[tid] = t;
    for (int m = 0; m < P.n_mod; ++m) {
        const int kind = P.mod[m].kind;
        if (!mgc_is_timeseries(kind)) continue;
        const int C = (kind == MGC_GRID) ? 4 : 1;
        S.base[m * MGC_STR + tid] = (int64_t)(P.series_nrm + (P.series_off[(int)V.cfg[P.mod[m].param_off]] + (int64_t)t * C));
    }
    const double *fstate = P.fstate + (int64_t)e * P.n_fstate;
    const int32_t *istate = P.istate + (int64_t)e * P.n_istate;
    for (int q = 0; q < P.n_nts; ++q) {
        const int32_t d = __ldg(P.nts + q);
        S.val[q * MGC_STR + tid] = mgc_obs_element(V, d >> 16, d & 0xffff, t, fstate, istate);
    }
}

// a series element: the windows of neighbouring envs and steps overlap -- keep their lines in L1 ahead of everything else
__device__ __forceinline__ double mgc_load_series(const double *p) {
    double v;
    asm volatile("ld.global.nc.L1::evict_last.f64 %0, [%1];" : "=d"(v) : "l"(p));
    return v;
}

// a row element: written once, never read by this kernel -- the store must not claim L1 lines the gathers live on
__device__ __forceinline__ void mgc_store_row(double *p, double v) {
#ifdef MGC_STORE_CS
    __stcs(p, v);
#else
    asm volatile("st.global.L1::no_allocate.f64 [%0], %1;" ::"l"(p), "d"(v) : "memory");
#endif
}

// one warp, rows r0, r0 + 4, ... of the tile.  NP > 0: the lane's table entries (elements lane, lane + 32, ... < 32 NP) are
// decoded ONCE, into the word of the staged tables the element starts from (a window base or a state value: the two
// tables are one array of 8-byte words) and its offset inside the window; a row whose whole horizon lies inside the series
// (t + hmax < T, one test per row) is then NP x (shared load, add, global load, streaming store) with nothing decoded.
// Rows that touch the end of the series, and rows longer than 256 elements (NP = 0: the table is re-read in blocks of
// four entries), decide per element.
template <int NP>
__device__ __forceinline__ void mgc_emit_rows_gather(const MgcLaunch &P, const MgcStage &S, double *obs, int e0, int n_tile,
                                                     int r0, int lane) {
    constexpr int NU = NP > 0 ? NP : 4;
    const int2 *tab = reinterpret_cast<const int2 *>(P.elem2);
    const int64_t *words = S.base;      // base[n_mod][STR] followed by val[n_nts][STR]
    int2 d[NU];
    int word[NU];
    uint32_t is_series = 0, active = 0;
    if (NP > 0) {
#pragma unroll
        for (int u = 0; u < NU; ++u) {
            const int j = lane + 32 * u;
            d[u] = (j < P.obs_dim) ? __ldg(tab + j) : make_int2(0, -1);
            if (d[u].x & 0x100) is_series |= 1u << u;
            if (d[u].x != 0 || d[u].y >= 0) active |= 1u << u;
            word[u] = (d[u].x & 0x100) ? (d[u].x & 0xff) * MGC_STR : (d[u].y >= 0 ? (P.n_mod + d[u].y) * MGC_STR : 0);
        }
    }
    for (int r = r0; r < n_tile; r += MGC_TILE / 32) {
        const int t = S.t[r];
        double *row = obs + (int64_t)(e0 + r) * P.obs_dim + lane;
        if (NP > 0 && t + P.hmax < P.T) {
            int64_t w[NU];
            double v[NU];
#pragma unroll
            for (int u = 0; u < NU; ++u) w[u] = words[word[u] + r];
#pragma unroll
            for (int u = 0; u < NU; ++u)
                v[u] = (is_series & (1u << u)) ? mgc_load_series(reinterpret_cast<const double *>(w[u]) + d[u].y) : __longlong_as_double(w[u]);
#pragma unroll
            for (int u = 0; u < NU; ++u)
                if (active & (1u << u)) mgc_store_row(row + 32 * u, v[u]);
            continue;
        }
        for (int j0 = 0; j0 < (NP > 0 ? 1 : P.obs_dim); j0 += 32 * NU) {
            double v[NU];
            uint32_t fill = 0;
            if (NP == 0) {
#pragma unroll
                for (int u = 0; u < NU; ++u) {
                    const int j = j0 + lane + 32 * u;
                    d[u] = (j < P.obs_dim) ? __ldg(tab + j) : make_int2(0, -1);
                }
            }
#pragma unroll
            for (int u = 0; u < NU; ++u) {
                v[u] = 0.0;
                if (d[u].x & 0x100) {
                    const int m = d[u].x & 0xff, h = d[u].x >> 9;
                    if (t < P.T && t + h < P.T) v[u] = __ldg(reinterpret_cast<const double *>(S.base[m * MGC_STR + r]) + d[u].y);
                    else fill |= 1u << u;
                } else if (d[u].y >= 0) {
                    v[u] = S.val[d[u].y * MGC_STR + r];
                }
            }
#pragma unroll
            for (int u = 0; u < NU; ++u) {
                if (fill & (1u << u))      // past the end of the series: the fill row, per element (rare)
                    v[u] = mgc_obs_element(mgc_view(P, e0 + r), d[u].x & 0xff, d[u].y, t, nullptr, nullptr);
                if (d[u].x != 0 || d[u].y >= 0) mgc_store_row(row + j0 + 32 * u, v[u]);
            }
        }
    }
}

// NP < 0: per-element decode (mgc_emit_row); otherwise the gather emitters above
template <int NP>
__global__ void __launch_bounds__(MGC_TILE) mgc_kernel(const __grid_constant__ MgcLaunch P) {
    extern __shared__ __align__(16) unsigned char mgc_smem[];
    constexpr bool kGather = NP >= 0;
    const int e0 = blockIdx.x * MGC_TILE;
    const int n_tile = min(MGC_TILE, P.n_envs - e0);
    const int e = e0 + threadIdx.x;
    const int n_steps = (P.mode == MGC_MODE_RUN || P.mode == MGC_MODE_RUN_DISCRETE) ? P.n_steps : 1;
    MgcStage S;
    if (kGather) S = mgc_stage(P, mgc_smem);
    for (int s = 0; s < n_steps; ++s) {
        if (threadIdx.x < n_tile) {
            mgc_owner(P, e, s);
            if (kGather && P.io.obs) mgc_stage_env(P, S, e, threadIdx.x);
        }
        __syncthreads();      // the tile's state rows (global memory, written by their owners) are read by every thread below
        if (P.io.obs) {
            double *obs = P.io.obs + (int64_t)(s % P.ring) * P.n_envs * P.obs_dim;
            if (kGather) {
                mgc_emit_rows_gather<(NP >= 0 ? NP : 0)>(P, S, obs, e0, n_tile, threadIdx.x >> 5, threadIdx.x & 31);
            } else {
                for (int r = threadIdx.x >> 5; r < n_tile; r += MGC_TILE / 32) mgc_emit_row(P, obs, e0 + r, threadIdx.x & 31);
            }
        }
        __syncthreads();      // the next step's owners overwrite the state this step's emitters have just read
    }
}

typedef void (*MgcKernel)(const MgcLaunch);
static MgcKernel mgc_gather_kernel_for(int obs_dim) {
    const int np = (obs_dim + 31) / 32;
    return np <= 2 ? mgc_kernel<2> : np <= 4 ? mgc_kernel<4> : np <= 8 ? mgc_kernel<8> : mgc_kernel<0>;
}
#else
static void mgc_kernel_host(const MgcLaunch &P) {
    const int tiles = (P.n_envs + MGC_TILE - 1) / MGC_TILE;
    const int n_steps = (P.mode == MGC_MODE_RUN || P.mode == MGC_MODE_RUN_DISCRETE) ? P.n_steps : 1;
    for (int b = 0; b < tiles; ++b) {
        const int e0 = b * MGC_TILE;
        const int n_tile = (P.n_envs - e0 < MGC_TILE) ? P.n_envs - e0 : MGC_TILE;
        for (int s = 0; s < n_steps; ++s) {
            for (int th = 0; th < n_tile; ++th) mgc_owner(P, e0 + th, s);
            if (P.io.obs) {
                double *obs = P.io.obs + (int64_t)(s % P.ring) * P.n_envs * P.obs_dim;
                for (int th = 0; th < MGC_TILE; ++th)
                    for (int r = th >> 5; r < n_tile; r += MGC_TILE / 32) mgc_emit_row(P, obs, e0 + r, th & 31);
            }
        }
    }
}
#endif

// forecast noise: a post-pass over rows that have just been written, one warp per row
struct MgcNoiseLaunch {
    const double *noise;
    double *obs;
    int64_t env_base;
    uint32_t k0, k1, c3, _pad;
};

MGC_DEV void mgc_noise_env(const MgcLaunch &P, const MgcNoiseLaunch &N, int e, int lane) {
    MgcView V = mgc_view(P, e);
    const double *sig = N.noise + (int64_t)P.cfg_index[e] * 2 * P.obs_dim;
    mgc_noise_row(V, P.elem, P.obs_dim, N.obs + (int64_t)e * P.obs_dim, sig, sig + P.obs_dim, P.step[e],
                  (uint64_t)(N.env_base + e), N.k0, N.k1, N.c3, lane);
}

#ifndef MGC_HOSTSIM
__global__ void __launch_bounds__(128) mgc_noise_kernel(const __grid_constant__ MgcLaunch P, const MgcNoiseLaunch N) {
    const int e = blockIdx.x * 4 + (threadIdx.x >> 5);
    if (e < P.n_envs) mgc_noise_env(P, N, e, threadIdx.x & 31);
}
#endif

// ---------------------------------------------------------------------------------------------------------------------
// C-ABI
// ---------------------------------------------------------------------------------------------------------------------
struct MgcHandle {
    MgcLaunch base;
    int64_t launches;
    int32_t *elem;      // the element table: device memory (host memory in the host build), owned by the handle
    int32_t *elem2, *nts;       // its gather form (device build only; see mgc_emit_row_gather)
    int gather_smem;            // dynamic shared memory of the gather kernel, 0 = the staged tables do not fit: decode per element
};

static int32_t *mgc_table_upload(const int32_t *host, int n) {
#ifdef MGC_HOSTSIM
    int32_t *p = new (std::nothrow) int32_t[n > 0 ? n : 1];
    if (p) memcpy(p, host, sizeof(int32_t) * (size_t)n);
    return p;
#else
    int32_t *p = nullptr;
    if (cudaMalloc(&p, sizeof(int32_t) * (size_t)(n > 0 ? n : 1)) != cudaSuccess) return nullptr;
    if (n > 0 && cudaMemcpy(p, host, sizeof(int32_t) * (size_t)n, cudaMemcpyHostToDevice) != cudaSuccess) {
        cudaFree(p);
        return nullptr;
    }
    return p;
#endif
}

static void mgc_table_free(int32_t *p) {
#ifdef MGC_HOSTSIM
    delete[] p;
#else
    if (p) {
        cudaFree(p);
        cudaGetLastError();     // a context that is already gone at interpreter exit is not an error worth keeping
    }
#endif
}

#ifdef MGC_HOSTSIM
static thread_local char g_mgc_err[512] = "";
extern "C" const char *mg_last_error(void) { return g_mgc_err; }
static int mgc_fail(int code, const char *msg) {
    snprintf(g_mgc_err, sizeof g_mgc_err, "%s", msg);
    return code;
}
#else
int mg_set_error(int code, const char *msg);      // mg_engine.cu: the text mg_last_error() returns
static int mgc_fail(int code, const char *msg) { return mg_set_error(code, msg); }
#endif

#ifdef MGC_HOSTSIM
// host build only: lets the CPU suite compare the streaming sum with numpy.sum directly, at every length of a growing list
extern "C" void mgc_test_growing_sums(const double *values, int n, double *sums) {
    MgcSum S;
    mgc_sum_init(S);
    sums[0] = mgc_sum_value(S);
    for (int i = 0; i < n; ++i) {
        mgc_sum_append(S, values[i]);
        sums[i + 1] = mgc_sum_value(S);
    }
}
#endif

extern "C" int mgc_abi_version(void) { return MGC_ABI_VERSION; }

extern "C" int64_t mgc_sizeof(int which) {
    switch (which) {
    case 0: return (int64_t)sizeof(MgcModule);
    case 1: return (int64_t)sizeof(MgcLayout);
    case 2: return (int64_t)sizeof(MgcIO);
    default: return -1;
    }
}

extern "C" int32_t mgc_param_count(int kind) {
    switch (kind) {
    case MGC_LOAD:
    case MGC_RENEWABLE: return 3;
    case MGC_BATTERY: return 6;
    case MGC_GENSET: return 8;
    case MGC_GRID: return 12;
    case MGC_UNBALANCED: return 2;
    default: return -1;
    }
}

extern "C" int mgc_create(const MgcLayout *L, MgcHandle **out) {
    if (!L || !out) return mgc_fail(MG_E_INVALID, "mgc_create: null argument");
    if (L->abi_version != MGC_ABI_VERSION) return mgc_fail(MG_E_INVALID, "mgc_create: abi_version mismatch");
    if (L->n_modules < 1 || L->n_modules > MGC_MAX_MODULES) return mgc_fail(MG_E_INVALID, "mgc_create: n_modules out of range");
    if (L->n_envs < 1 || L->n_envs >= (1ll << 31) - MGC_TILE) return mgc_fail(MG_E_INVALID, "mgc_create: n_envs out of range");
    if (L->n_cfg < 1 || L->cfg_stride < MGC_CFG_HEADER || !L->cfg || !L->step || !L->cfg_index)
        return mgc_fail(MG_E_INVALID, "mgc_create: missing config / state arrays");
    // the module table: dispatch classes in order (fixed, controllable, flex), blocks inside the rows, listing a permutation
    int cls = 0, obs = 0, act = 0, nf = 0, ni = 0, n_ts = 0;
    bool seen[MGC_MAX_MODULES] = {false};
    for (int m = 0; m < L->n_modules; ++m) {
        const MgcModule &M = L->modules[m];
        const int pc = mgc_param_count(M.kind);
        if (pc < 0) return mgc_fail(MG_E_INVALID, "mgc_create: unknown module kind");
        const int c = mgc_dispatch_class(M.kind);
        if (c < cls) return mgc_fail(MG_E_INVALID, "mgc_create: modules must be in dispatch order (fixed, controllable, flex)");
        cls = c;
        if (M.param_off < MGC_CFG_HEADER || M.param_off + pc > L->cfg_stride)
            return mgc_fail(MG_E_INVALID, "mgc_create: module parameters outside the config record");
        if (M.horizon < 0 || (!mgc_is_timeseries(M.kind) && M.horizon != 0)) return mgc_fail(MG_E_INVALID, "mgc_create: bad horizon");
        const int len = mgc_obs_len(M);
        if (!L->obs_select && len > 0 && (M.obs_off < 0 || M.obs_off + len > L->obs_dim))
            return mgc_fail(MG_E_INVALID, "mgc_create: observation block outside the row");
        obs += len;
        if (c == 1) {
            const int w = (M.kind == MGC_GENSET) ? 2 : 1;
            if (M.act_col < 0 || M.act_col + w > L->n_act) return mgc_fail(MG_E_INVALID, "mgc_create: action columns outside the row");
            act += w;
        }
        if (M.kind == MGC_BATTERY) {
            if (M.fstate_off < 0 || M.fstate_off + 2 > L->n_fstate) return mgc_fail(MG_E_INVALID, "mgc_create: battery state outside the row");
            nf += 2;
        }
        if (M.kind == MGC_GENSET) {
            if (M.istate_off < 0 || M.istate_off + 4 > L->n_istate) return mgc_fail(MG_E_INVALID, "mgc_create: genset state outside the row");
            ni += 4;
        }
        if (mgc_is_timeseries(M.kind)) ++n_ts;
        if (M.listing < 0 || M.listing >= L->n_modules || seen[M.listing]) return mgc_fail(MG_E_INVALID, "mgc_create: listing is not a permutation");
        seen[M.listing] = true;
    }
    if (L->obs_select) obs = L->obs_dim;      // the row holds the selected elements only (checked below)
    if (L->obs_dim < 0 || obs != L->obs_dim || act != L->n_act || nf != L->n_fstate || ni != L->n_istate)
        return mgc_fail(MG_E_INVALID, "mgc_create: row widths do not match the module table");
    if ((nf > 0 && !L->fstate) || (ni > 0 && !L->istate)) return mgc_fail(MG_E_INVALID, "mgc_create: null state pointer");
    if (n_ts > 0 && (L->series_len < 1 || L->n_series < 1 || !L->series || !L->series_off))
        return mgc_fail(MG_E_INVALID, "mgc_create: time-series modules without a series pool");
    if (L->plist && (L->n_plist < 1 || L->plist_width < 1 || L->plist_width > 2 * MGC_MAX_MODULES))
        return mgc_fail(MG_E_INVALID, "mgc_create: bad priority-list table");
    // element -> (module, offset): every element of the row belongs to exactly one module block
    int32_t *tab = new (std::nothrow) int32_t[L->obs_dim > 0 ? L->obs_dim : 1];
    if (!tab) return mgc_fail(MG_E_INVALID, "mgc_create: out of host memory");
    for (int j = 0; j < L->obs_dim; ++j) tab[j] = -1;
    if (L->obs_select) {
        for (int j = 0; j < L->obs_dim; ++j) {
            const int32_t d = L->obs_select[j];
            const int m = d >> 16, k = d & 0xffff;
            if (d < 0 || m >= L->n_modules || k >= mgc_obs_len(L->modules[m])) {
                delete[] tab;
                return mgc_fail(MG_E_INVALID, "mgc_create: obs_select entry outside the module's observation block");
            }
            tab[j] = d;
        }
    }
    for (int m = 0; m < L->n_modules && !L->obs_select; ++m) {
        const int len = mgc_obs_len(L->modules[m]);
        if (len > 0xffff) { delete[] tab; return mgc_fail(MG_E_UNSUPPORTED, "mgc_create: a module's observation block exceeds 65535 elements"); }
        for (int k = 0; k < len; ++k) {
            int32_t &slot = tab[L->modules[m].obs_off + k];
            if (slot != -1) { delete[] tab; return mgc_fail(MG_E_INVALID, "mgc_create: observation blocks overlap"); }
            slot = (m << 16) | k;
        }
    }
    MgcHandle *h = new (std::nothrow) MgcHandle();
    if (!h) { delete[] tab; return mgc_fail(MG_E_INVALID, "mgc_create: out of host memory"); }
    h->elem = mgc_table_upload(tab, L->obs_dim);
    h->elem2 = h->nts = nullptr;
    h->gather_smem = 0;
    int n_nts = 0, hmax_row = 0;
#ifndef MGC_HOSTSIM
    const char *gather_env = getenv("PYMGRID_B200_COMPOSE_GATHER");      // "0": keep the per-element decode (A/B and tests)
    if (h->elem && L->series_nrm && L->obs_dim > 0 && !(gather_env && gather_env[0] == '0')) {
        // gather form: series elements carry (module, forecast row, offset), the others index the list of state elements
        int32_t *tab2 = new (std::nothrow) int32_t[3 * (size_t)L->obs_dim];
        if (!tab2) { delete[] tab; mgc_table_free(h->elem); delete h; return mgc_fail(MG_E_INVALID, "mgc_create: out of host memory"); }
        int32_t *list = tab2 + 2 * (size_t)L->obs_dim;
        bool fits = true;
        int hmax = 0;
        for (int j = 0; j < L->obs_dim; ++j) {
            const int m = tab[j] >> 16, k = tab[j] & 0xffff;
            const int kind = L->modules[m].kind;
            if (mgc_is_timeseries(kind)) {
                const int row = k / ((kind == MGC_GRID) ? 4 : 1);
                if (row >= (1 << 22)) fits = false;
                if (row > hmax) hmax = row;
                tab2[2 * j] = m | 0x100 | (row << 9);
                tab2[2 * j + 1] = k;
            } else {
                tab2[2 * j] = 0;
                tab2[2 * j + 1] = n_nts;
                list[n_nts++] = tab[j];
            }
        }
        MgcLaunch probe;
        probe.n_mod = L->n_modules;
        probe.n_nts = n_nts;
        const size_t bytes = mgc_stage_bytes(probe);
        if (fits && bytes <= 96 * 1024) {
            h->elem2 = mgc_table_upload(tab2, 2 * L->obs_dim);
            h->nts = mgc_table_upload(list, n_nts);
            if (h->elem2 && h->nts &&
                cudaFuncSetAttribute(mgc_gather_kernel_for(L->obs_dim), cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024) == cudaSuccess) {
                h->gather_smem = (int)bytes;
            } else {
                cudaGetLastError();
                mgc_table_free(h->elem2);
                mgc_table_free(h->nts);
                h->elem2 = h->nts = nullptr;
            }
        }
        hmax_row = hmax;
        delete[] tab2;
    }
#endif
    delete[] tab;
    if (!h->elem) { delete h; return mgc_fail(MG_E_CUDA, "mgc_create: could not allocate the element table on the device"); }
    MgcLaunch &B = h->base;
    memset(&B, 0, sizeof B);
    B.elem = h->elem;
    B.elem2 = h->elem2; B.nts = h->nts; B.n_nts = n_nts; B.gather = h->gather_smem > 0; B.hmax = hmax_row;
    for (int m = 0; m < L->n_modules; ++m)
        B.n_act_modules += (L->modules[m].kind == MGC_GENSET) ? 2 : (L->modules[m].kind == MGC_LOAD) ? 0 : 1;
    memcpy(B.mod, L->modules, sizeof(MgcModule) * (size_t)L->n_modules);
    B.n_mod = L->n_modules; B.n_act = L->n_act; B.obs_dim = L->obs_dim; B.n_fstate = L->n_fstate; B.n_istate = L->n_istate;
    B.cfg_stride = L->cfg_stride; B.T = L->series_len; B.n_envs = (int32_t)L->n_envs;
    B.cfg = L->cfg; B.series = L->series; B.series_off = L->series_off; B.series_nrm = L->series_nrm;
    B.step = L->step; B.fstate = L->fstate; B.istate = L->istate; B.cfg_index = L->cfg_index;
    B.plist = L->plist; B.n_plist = L->plist ? L->n_plist : 0; B.plist_width = L->plist_width;
    B.env_initial = L->env_initial_step; B.env_final = L->env_final_step;
    h->launches = 0;
    *out = h;
    return MG_OK;
}

extern "C" int mgc_destroy(MgcHandle *h) {
    if (h) {
        mgc_table_free(h->elem);
        mgc_table_free(h->elem2);
        mgc_table_free(h->nts);
    }
    delete h;
    return MG_OK;
}

extern "C" int64_t mgc_launch_count(const MgcHandle *h) { return h ? h->launches : 0; }

static int mgc_launch(MgcHandle *h, const MgcIO *io, int mode, int32_t n_steps, int32_t ring, int normalized, void *stream) {
    if (!h || !io) return mgc_fail(MG_E_INVALID, "mgc: null argument");
    MgcLaunch P = h->base;
    P.io = *io;
    P.mode = mode; P.n_steps = n_steps; P.ring = ring; P.normalized = normalized;
    if (mode == MGC_MODE_RUN) {
        if (n_steps < 1 || ring < 1) return mgc_fail(MG_E_INVALID, "mgc_run: n_steps and ring must be >= 1");
        if (!io->reward || !io->done) return mgc_fail(MG_E_INVALID, "mgc_run: reward and done are required");
        if (P.n_act > 0 && !io->actions) return mgc_fail(MG_E_INVALID, "mgc_run: actions are required (n_act > 0)");
    } else if (mode == MGC_MODE_MODULES) {
        if (!io->reward || !io->done) return mgc_fail(MG_E_INVALID, "mgc_modules_step: reward and done are required");
        if (P.n_act_modules > 0 && !io->actions) return mgc_fail(MG_E_INVALID, "mgc_modules_step: actions are required");
    } else if (mode == MGC_MODE_RUN_DISCRETE) {
        if (n_steps < 1 || ring < 1) return mgc_fail(MG_E_INVALID, "mgc_run_discrete: n_steps and ring must be >= 1");
        if (!io->reward || !io->done || !io->dactions) return mgc_fail(MG_E_INVALID, "mgc_run_discrete: reward, done and dactions are required");
        if (!P.plist) return mgc_fail(MG_E_INVALID, "mgc_run_discrete: the layout has no priority-list table");
    } else {
        P.n_steps = 1; P.ring = 1;
        if (mode == MGC_MODE_OBSERVE && !io->obs) return mgc_fail(MG_E_INVALID, "mgc_observe: obs is required");
    }
#ifdef MGC_HOSTSIM
    (void)stream;
    mgc_kernel_host(P);
#else
    const int tiles = (P.n_envs + MGC_TILE - 1) / MGC_TILE;
    if (P.gather) mgc_gather_kernel_for(P.obs_dim)<<<tiles, MGC_TILE, h->gather_smem, (cudaStream_t)stream>>>(P);
    else mgc_kernel<-1><<<tiles, MGC_TILE, 0, (cudaStream_t)stream>>>(P);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) {
        char msg[256];
        snprintf(msg, sizeof msg, "mgc kernel launch: %s", cudaGetErrorString(e));
        return mgc_fail(MG_E_CUDA, msg);
    }
#endif
    h->launches += 1;
    return MG_OK;
}

extern "C" int mgc_run(MgcHandle *h, const MgcIO *io, int32_t n_steps, int32_t ring, int normalized, void *stream) {
    return mgc_launch(h, io, MGC_MODE_RUN, n_steps, ring, normalized, stream);
}
extern "C" int mgc_run_discrete(MgcHandle *h, const MgcIO *io, int32_t n_steps, int32_t ring, void *stream) {
    return mgc_launch(h, io, MGC_MODE_RUN_DISCRETE, n_steps, ring, 0, stream);
}
extern "C" int mgc_modules_step(MgcHandle *h, const MgcIO *io, int normalized, void *stream) {
    return mgc_launch(h, io, MGC_MODE_MODULES, 1, 1, normalized, stream);
}
extern "C" int mgc_forecast_noise(MgcHandle *h, const double *noise, double *obs, int64_t env_base, uint64_t seed, uint64_t call,
                                  void *stream) {
    if (!h || !noise || !obs) return mgc_fail(MG_E_INVALID, "mgc_forecast_noise: null argument");
    if (h->base.obs_dim > 2 * 65535) return mgc_fail(MG_E_UNSUPPORTED, "mgc_forecast_noise: observation row too long");
    MgcNoiseLaunch N;
    N.noise = noise; N.obs = obs; N.env_base = env_base;
    N.k0 = (uint32_t)seed;
    N.k1 = (uint32_t)(seed >> 32) ^ (uint32_t)(call >> 32);
    N.c3 = (uint32_t)call;
    N._pad = 0;
#ifdef MGC_HOSTSIM
    (void)stream;
    for (int e = 0; e < h->base.n_envs; ++e)
        for (int lane = 0; lane < 32; ++lane) mgc_noise_env(h->base, N, e, lane);
#else
    mgc_noise_kernel<<<(h->base.n_envs + 3) / 4, 128, 0, (cudaStream_t)stream>>>(h->base, N);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) {
        char msg[256];
        snprintf(msg, sizeof msg, "mgc_forecast_noise launch: %s", cudaGetErrorString(e));
        return mgc_fail(MG_E_CUDA, msg);
    }
#endif
    h->launches += 1;
    return MG_OK;
}
extern "C" int mgc_reset(MgcHandle *h, const MgcIO *io, void *stream) { return mgc_launch(h, io, MGC_MODE_RESET, 1, 1, 0, stream); }
extern "C" int mgc_observe(MgcHandle *h, const MgcIO *io, void *stream) { return mgc_launch(h, io, MGC_MODE_OBSERVE, 1, 1, 0, stream); }
