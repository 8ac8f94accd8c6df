// d3q19_api.cu -- the C-ABI of include/d3q19_b200.h: handle, state transfer, step
// orchestration (split boundary/interior launches, NCCL z-face exchange on a second stream),
// macrovar / rhoupdat / avedensity / probe / profiles, and the shim state machine that the
// replacement collision.f90 calls.  There is no CPU path in this file: every entry point
// that computes launches a kernel from kernels.cuh on the handle's GPU.
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <new>
#include <string>
#include <vector>

#include "../../include/d3q19_b200.h"
#include "kernels.cuh"
#include "particles.cuh"
#include "nccl_dl.h"

using namespace d3q;

// ---- error plumbing ------------------------------------------------------------------------
static thread_local std::string g_err;

static int fail(const char *fmt, ...) {
    char buf[1024];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    g_err = buf;
    return 1;
}

#define CK(call)                                                                                   \
    do {                                                                                           \
        cudaError_t e_ = (call);                                                                   \
        if (e_ != cudaSuccess) return fail("%s:%d %s -> %s", __FILE__, __LINE__, #call, cudaGetErrorString(e_)); \
    } while (0)
#define NK(call)                                                                                   \
    do {                                                                                           \
        int e_ = (call);                                                                           \
        if (e_ != 0) return fail("%s:%d %s -> NCCL %s", __FILE__, __LINE__, #call, nccl_api().GetErrorString(e_)); \
    } while (0)
#define RK_(call)                                 \
    do {                                          \
        int e_ = (call);                          \
        if (e_ != 0) return e_;                   \
    } while (0)

extern "C" const char *d3q19_last_error(void) { return g_err.c_str(); }

// ---- the handle --------------------------------------------------------------------------------
struct Shim {
    d3q19_shim_arrays a;
    bool bound = false;
    bool f_dev_valid = false;     // device populations reflect the latest state
    bool f_host_valid = false;    // host f(:,:,:,:) reflects the latest state
    bool macro_dev_valid = false; // device rho,u = moments of the current f (or what the driver set)
    bool u_dev_loaded = false;    // host ux,uy,uz were uploaded (frozen-u pre-relaxation)
    int next_mode = D3Q19_MACRO_MAIN;
    bool collided_since_macrovar = false;  // a main-loop collision_MRT ran since the last macrovar
    bool in_prerelax = false;     // the last rhoupdat has not been followed by its collision_MRT yet
    int prerelax_iter = 0;        // main.f90's istep inside the pre-relaxation loop
    double prerelax_err = 0.0;    // max|rho - rhop| over all ranks of the current iteration (main.f90:79-80)
    void *pinned[5] = {nullptr, nullptr, nullptr, nullptr, nullptr};   // host arrays this library page-locked (shim_pin)
    int npinned = 0;
};

struct d3q19_handle {
    d3q19_config cfg;
    Geom g;
    size_t nfield = 0;            // xp*ly*lz
    double *A = nullptr, *B = nullptr;
    double *A_alloc = nullptr, *B_alloc = nullptr;   // A, B sit PAD elements inside these (x-1 / x+1 past the ends stays mapped)
    bool idx32 = true;            // a population has < 2^32 elements: 32-bit in-slab indices
    int phase = 0;                // AA: 0 canonical, 1 swapped post-collision.  AB: always post-collision.
    double *rho = nullptr, *ux = nullptr, *uy = nullptr, *uz = nullptr;
    double *ffx = nullptr, *ffy = nullptr, *ffz = nullptr;
    double *vort = nullptr;       // ox, oy, oz: 3 x nfield (vortcalc)
    double *vort_halo = nullptr;  // exchanged velocity planes: [lo: ux,uy,uz][hi: ux,uy,uz]
    double *sij2 = nullptr;       // Sij*Sij, nfield (sijstat)
    int32_t *solid = nullptr, *isn = nullptr;
    double *ypglb = nullptr, *wp = nullptr, *omgp = nullptr;
    int npart = 0;
    double Fx = 0, Fy = 0, Fz = 0, rho_shift = 0;
    cudaStream_t sc = nullptr, sx = nullptr;
    cudaEvent_t evB = nullptr, evX = nullptr, t0 = nullptr, t1 = nullptr, evC[2] = {nullptr, nullptr},
                evS[2] = {nullptr, nullptr};
    bool exchange_pending = false;
    // optional per-step timeline (d3q19_trace_enable): 4 timing events per step -- [0] before the boundary launch,
    // [1] after it, [2] after the interior launch (all on sc), [3] after the exchange / put (on sx)
    cudaEvent_t *trace_ev = nullptr;
    int trace_cap = 0, trace_n = 0;
    bool put_pending = false;     // copy-engine transport: the last copies (on sx, event evX) may still be reading my planes
    NcclComm comm = nullptr;
    double *send_up = nullptr, *send_dn = nullptr, *recv_lo = nullptr, *recv_hi = nullptr;
    double *stage[2] = {nullptr, nullptr};
    int stage_planes = 0;
    double *scal = nullptr;       // small device scratch (64 doubles)
    double *red_d = nullptr;      // reduction partials
    long long *red_c = nullptr;
    double *prof_partial = nullptr, *prof_out = nullptr;
    double *diag_partial = nullptr, *diag_red = nullptr;   // d3q19_diag scratch (allocated once: no cudaMalloc/cudaFree in the loop)
    int diag_npartial = 0;
    int prof_chunks = 0, prof_rows = 0;
    long long n_step_kernels = 0, n_other_kernels = 0, n_nccl = 0, n_steps = 0, n_copies = 0;
    // particle path (particles.cuh)
    bool part_on = false, links_valid = false, mask_built = false;
    d3q19_particle_params pp;
    int32_t *own = nullptr;                      // ghosted owner mask [lz+2][ly][xp]: particles.cuh "the solid mask follows the particles"
    double *pbuf = nullptr;                      // 10 tables of (3,npart)
    double *ypmask = nullptr,                    // the positions the mask was built with
           *fHIp = nullptr, *torqp = nullptr, *flubp = nullptr, *forcepp = nullptr, *torqpp = nullptr, *thetap = nullptr;
    Links links = {nullptr, nullptr, nullptr, 0};
    FillList fill = {nullptr, nullptr, nullptr, 0};   // nodes the last mask update uncovered (what beads_filling rebuilds)
    long long maxlink = 0, nlink = 0;
    unsigned long long *pcnt = nullptr;          // device counters: [1] refill list, [2] refilled nodes (the link counts live in links.count)
    int part_rows = 0;                           // rows of the largest bounding box (grid of the sweep kernels)
    double *fill_halo = nullptr;                 // z-slab refill: [send up][send dn][ghost lo][ghost hi], 19 x plane each
    cudaEvent_t evP = nullptr, evF = nullptr;    // refill source exchange on sx next to the bookkeeping kernels on sc
    bool fill_xchg_pending = false;
    double amp = 0, aip = 0;
    // halo in peer memory (cudaIpc): [0] = lower neighbour (mzm), [1] = upper neighbour (mzp)
    bool halo_on = false;
    int halo_mode = D3Q19_HALO_FUSED;            // D3Q19_HALO_FUSED: stores inside the step kernel; D3Q19_HALO_PUT: a copy kernel on sx
    unsigned int halo_epoch = 0;
    unsigned int *halo_flags = nullptr;          // local: [0] wait_lo, [1] wait_hi, [2..3] block counters, [8] watchdog
    int halo_split_min = 64;      // peer-memory halo: slabs at least this thick run boundary and interior as two launches
    long long pf_ahead = 0;       // AA steps: L2 prefetch distance in elements (kernels.cuh "Software prefetch")
    unsigned long long halo_timeout_ns = 30ull * 1000000000ull;   // D3Q19_HALO_TIMEOUT_S
    void *peer_base[2][3] = {{nullptr, nullptr, nullptr}, {nullptr, nullptr, nullptr}};   // opened A_alloc, B_alloc, flags
    double *peer_A[2] = {nullptr, nullptr}, *peer_B[2] = {nullptr, nullptr};
    unsigned int *peer_flags[2] = {nullptr, nullptr};
    long long peer_slab[2] = {0, 0};
    int peer_lz[2] = {0, 0};
    Shim shim;
};

static void shim_unpin(d3q19_handle *h);

static const size_t POP_PAD = 32;     // doubles (256 B) in front of and behind the populations

static const FaceSlots SLOTS_PZ = {{5, 11, 12, 15, 16}};   // c_z = +1 (collision.f90:337-341)
static const FaceSlots SLOTS_MZ = {{6, 13, 14, 17, 18}};   // c_z = -1 (collision.f90:343-347)

static inline dim3 grid_nodes(const d3q19_handle *h, int nplanes) {
    return dim3((unsigned)((h->g.lx + BLOCK_X - 1) / BLOCK_X), (unsigned)h->g.ly, (unsigned)nplanes);
}

static int read_kind(const d3q19_handle *h) {
    if (h->cfg.scheme == D3Q19_SCHEME_AB) return READ_PULL_NAT;
    return h->phase == 0 ? READ_DIRECT : READ_PULL_SWAP;
}

static int ensure_macro_arrays(d3q19_handle *h) {
    if (h->rho) return 0;
    const size_t bytes = h->nfield * sizeof(double);
    CK(cudaMalloc(&h->rho, bytes)); CK(cudaMalloc(&h->ux, bytes));
    CK(cudaMalloc(&h->uy, bytes)); CK(cudaMalloc(&h->uz, bytes));
    CK(cudaMemsetAsync(h->rho, 0, bytes, h->sc)); CK(cudaMemsetAsync(h->ux, 0, bytes, h->sc));
    CK(cudaMemsetAsync(h->uy, 0, bytes, h->sc)); CK(cudaMemsetAsync(h->uz, 0, bytes, h->sc));
    return 0;
}

static int ensure_stage(d3q19_handle *h) {
    if (h->stage[0]) return 0;
    const size_t per_plane = (size_t)NPOP * h->g.lx * h->g.ly * sizeof(double);
    size_t planes = (size_t)(192u << 20) / per_plane;
    if (planes < 1) planes = 1;
    if (planes > (size_t)h->g.lz) planes = (size_t)h->g.lz;
    h->stage_planes = (int)planes;
    CK(cudaMalloc(&h->stage[0], planes * per_plane));
    CK(cudaMalloc(&h->stage[1], planes * per_plane));
    return 0;
}

// sc must not read ghost/boundary data before the last exchange finished
static int wait_exchange(d3q19_handle *h) {
    if (h->halo_on && h->halo_epoch > 0) {
        // the neighbours' stores of the last halo step must have landed before anybody reads the planes
        k_halo_wait<<<1, 1, 0, h->sc>>>(h->halo_flags, h->halo_flags + 1, h->halo_epoch, h->halo_flags + 8, h->halo_timeout_ns);
        CK(cudaGetLastError());
    }
    if (h->exchange_pending || h->put_pending) {
        CK(cudaStreamWaitEvent(h->sc, h->evX, 0));
        h->exchange_pending = false;
        h->put_pending = false;
    }
    return 0;
}

// particle runs: whoever reads the solid mask (the step, macrovar, avedensity, diag, the plane sums) must see the mask of
// the CURRENT particle table; d3q19_beads_links is a no-op while the table has not changed since the last build
extern "C" int d3q19_beads_links(d3q19_handle *h, int64_t *nlink_local);
static int check_link_overflow(d3q19_handle *h, const char *who);
static int ensure_mask(d3q19_handle *h) {
    if (h->part_on && !h->links_valid) return d3q19_beads_links(h, nullptr);
    return 0;
}

static inline void trace_mark(d3q19_handle *h, int which, cudaStream_t s) {
    if (h->trace_ev && h->trace_n < h->trace_cap) cudaEventRecord(h->trace_ev[4 * h->trace_n + which], s);
}
static inline void trace_next(d3q19_handle *h) {
    if (h->trace_ev && h->trace_n < h->trace_cap) h->trace_n++;
}

// ---- z-face exchange (collisionExchnge's z phase, collision.f90:337-370) --------------------------
// "up" data goes to the +z neighbour, which stores it at plane lo_dst; "dn" data goes to the
// -z neighbour, which stores it at plane hi_dst.  Runs on h->sx after `after` (an event on sc).
static int exchange_faces(d3q19_handle *h, double *arr, int up_src, const FaceSlots &up_slots, int lo_dst,
                          int dn_src, const FaceSlots &dn_slots, int hi_dst, int exclude_walls, cudaStream_t s) {
    const Geom &g = h->g;
    const size_t cnt = (size_t)5 * g.plane;
    const int up = (h->cfg.rank + 1) % h->cfg.nranks;                       // mzp, para.f90:266
    const int dn = (h->cfg.rank + h->cfg.nranks - 1) % h->cfg.nranks;       // mzm, para.f90:267
    const dim3 gp((unsigned)((g.xp + BLOCK_X - 1) / BLOCK_X), (unsigned)g.ly, 10u);
    FacePair pk;
    pk.buf[0] = h->send_up; pk.zg[0] = up_src; pk.slots[0] = up_slots;
    pk.buf[1] = h->send_dn; pk.zg[1] = dn_src; pk.slots[1] = dn_slots;
    k_face_pack<<<gp, BLOCK_X, 0, s>>>(g, arr, pk);
    CK(cudaGetLastError());
    NcclApi &n = nccl_api();
    NK(n.GroupStart());
    NK(n.Send(h->send_up, cnt, NCCL_FLOAT64, up, h->comm, s));
    NK(n.Send(h->send_dn, cnt, NCCL_FLOAT64, dn, h->comm, s));
    NK(n.Recv(h->recv_lo, cnt, NCCL_FLOAT64, dn, h->comm, s));
    NK(n.Recv(h->recv_hi, cnt, NCCL_FLOAT64, up, h->comm, s));
    NK(n.GroupEnd());
    const dim3 gu((unsigned)((g.lx + BLOCK_X - 1) / BLOCK_X), (unsigned)g.ly, 10u);
    FacePair un;
    un.buf[0] = h->recv_lo; un.zg[0] = lo_dst; un.slots[0] = up_slots;
    un.buf[1] = h->recv_hi; un.zg[1] = hi_dst; un.slots[1] = dn_slots;
    k_face_unpack<<<gu, BLOCK_X, 0, s>>>(g, arr, un, exclude_walls);
    CK(cudaGetLastError());
    h->n_other_kernels += 2;
    h->n_nccl += 4;
    return 0;
}

// the exchange that must follow a step of the given kind
static int exchange_after_step(d3q19_handle *h, int step_kind, double *arr, cudaStream_t s) {
    const int lz = h->g.lz;
    switch (step_kind) {
    case STEP_AB:       // post-collision, natural slots: my plane lz (c_z=+1 slots) -> +z ghost 0
        return exchange_faces(h, arr, lz, SLOTS_PZ, 0, 1, SLOTS_MZ, lz + 1, 0, s);
    case STEP_AA_EVEN:  // post-collision, swapped slots: f*_i sits in slot opp(i)
        return exchange_faces(h, arr, lz, SLOTS_MZ, 0, 1, SLOTS_PZ, lz + 1, 0, s);
    default:            // AA odd pushed across the faces into my ghosts: ghost -> neighbour's real plane
        return exchange_faces(h, arr, lz + 1, SLOTS_PZ, 1, 0, SLOTS_MZ, lz, 1, s);
    }
}

// ---- life cycle ----------------------------------------------------------------------------------
extern "C" int d3q19_device_count(int32_t *n) {
    int c = 0;
    cudaError_t e = cudaGetDeviceCount(&c);
    if (e != cudaSuccess) { *n = 0; return fail("cudaGetDeviceCount -> %s", cudaGetErrorString(e)); }
    *n = c;
    return 0;
}

extern "C" int d3q19_nccl_unique_id(unsigned char out[128]) {
    NcclApi &n = nccl_api();
    if (const char *m = n.load()) return fail("%s", m);
    NcclUniqueId id;
    NK(n.GetUniqueId(&id));
    memcpy(out, id.internal, 128);
    return 0;
}

extern "C" int d3q19_destroy(d3q19_handle *h) {
    if (!h) return 0;
    cudaSetDevice(h->cfg.device);
    if (h->sc) cudaStreamSynchronize(h->sc);
    if (h->sx) cudaStreamSynchronize(h->sx);
    shim_unpin(h);
    if (h->halo_on) {
        // nobody may still be storing into our planes, and we must let go of theirs before they free them
        for (int d = 0; d < 2; ++d)
            for (int a = 0; a < 3; ++a)
                if (h->peer_base[d][a] && !(d == 1 && h->peer_base[0][a] == h->peer_base[1][a])) cudaIpcCloseMemHandle(h->peer_base[d][a]);
        if (h->comm && h->scal) {
            NcclApi &n = nccl_api();
            n.AllReduce(h->scal + 32, h->scal + 32, 1, NCCL_FLOAT64, NCCL_SUM, h->comm, h->sc);
            cudaStreamSynchronize(h->sc);
        }
        h->halo_on = false;
    }
    if (h->part_on) {
        void *pp_[] = {h->own, h->pbuf, h->links.node, h->links.dir, h->links.count, h->fill.node, h->fill.part,
                       h->pcnt, h->fill_halo};
        for (void *q : pp_) if (q) cudaFree(q);
        h->solid = h->isn = nullptr; h->ypglb = h->wp = h->omgp = nullptr;
    }
    if (h->halo_flags) cudaFree(h->halo_flags);
    if (h->trace_ev) {
        for (int i = 0; i < 4 * h->trace_cap; ++i) cudaEventDestroy(h->trace_ev[i]);
        delete[] h->trace_ev;
    }
    if (h->comm) nccl_api().CommDestroy(h->comm);
    void *ptrs[] = {h->A_alloc, h->B_alloc, h->rho, h->ux, h->uy, h->uz, h->ffx, h->ffy, h->ffz, h->solid, h->isn, h->ypglb,
                    h->wp, h->omgp, h->send_up, h->send_dn, h->recv_lo, h->recv_hi, h->stage[0], h->stage[1],
                    h->scal, h->red_d, h->red_c, h->prof_partial, h->prof_out, h->vort, h->vort_halo, h->diag_partial, h->diag_red, h->sij2};
    for (void *p : ptrs) if (p) cudaFree(p);
    cudaEvent_t evs[] = {h->evP, h->evF, h->evB, h->evX, h->t0, h->t1, h->evC[0], h->evC[1], h->evS[0], h->evS[1]};
    for (cudaEvent_t e : evs) if (e) cudaEventDestroy(e);
    if (h->sc) cudaStreamDestroy(h->sc);
    if (h->sx) cudaStreamDestroy(h->sx);
    delete h;
    return 0;
}

extern "C" int d3q19_create(const d3q19_config *cfg, d3q19_handle **out) {
    if (!cfg || !out) return fail("d3q19_create: null argument");
    *out = nullptr;
    if (cfg->abi_version != D3Q19_ABI_VERSION)
        return fail("d3q19_create: abi_version %d, library speaks %d", cfg->abi_version, D3Q19_ABI_VERSION);
    if (cfg->lx < 2 || cfg->ly < 1 || cfg->lz < 1) return fail("d3q19_create: bad local extents %d %d %d", cfg->lx, cfg->ly, cfg->lz);
    if (cfg->nranks < 1 || cfg->rank < 0 || cfg->rank >= cfg->nranks) return fail("d3q19_create: bad rank %d of %d", cfg->rank, cfg->nranks);
    if (cfg->scheme != D3Q19_SCHEME_AA && cfg->scheme != D3Q19_SCHEME_AB && cfg->scheme != D3Q19_SCHEME_AUTO)
        return fail("d3q19_create: unknown scheme %d", cfg->scheme);
    if (cfg->ly > 65535 || cfg->lz + 2 > 65535) return fail("d3q19_create: ly, lz must be < 65535");
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev < 1)
        return fail("d3q19_create: no CUDA device -- this library has no CPU fallback");
    if (cfg->device < 0 || cfg->device >= ndev) return fail("d3q19_create: device %d of %d", cfg->device, ndev);
    CK(cudaSetDevice(cfg->device));

    d3q19_handle *h = new (std::nothrow) d3q19_handle();
    if (!h) return fail("d3q19_create: out of host memory");
    h->cfg = *cfg;
    Geom &g = h->g;
    g.lx = cfg->lx; g.ly = cfg->ly; g.lz = cfg->lz;
    g.xp = (cfg->lx + 15) / 16 * 16;
    g.plane = (long long)g.xp * g.ly;
    g.slab = g.plane * (g.lz + 2);
    g.zlo_src = cfg->nranks == 1 ? g.lz : 0;
    g.zhi_src = cfg->nranks == 1 ? 1 : g.lz + 1;
    h->nfield = (size_t)g.plane * g.lz;
    {
        // prefetch about 128 thread blocks ahead of the running front, in whole x-rows (measured optimum on
        // B200: profiles/r01c_prefetch_sweep.md); d3q19_config.pf_blocks overrides, < 0 switches it off
        const int pf_blocks = cfg->pf_blocks == 0 ? 128 : (cfg->pf_blocks < 0 ? 0 : cfg->pf_blocks);
        const int blocks_per_row = (g.lx + BLOCK_X - 1) / BLOCK_X;
        const long long rows = pf_blocks > 0 ? (pf_blocks + blocks_per_row - 1) / blocks_per_row : 0;
        h->pf_ahead = rows * g.xp;
    }
    if (h->cfg.scheme == D3Q19_SCHEME_AUTO) {
        // two arrays (one-step pull) when they fit comfortably, else in place
        size_t free_b = 0, total_b = 0;
        if (cudaMemGetInfo(&free_b, &total_b) != cudaSuccess) { fail("d3q19_create: cudaMemGetInfo failed"); delete h; return 1; }
        const double two = 2.0 * NPOP * (double)g.slab * sizeof(double);
        h->cfg.scheme = (two < 0.45 * (double)free_b) ? D3Q19_SCHEME_AB : D3Q19_SCHEME_AA;
    }

#define CKH(call)                                                      \
    do {                                                               \
        cudaError_t e_ = (call);                                       \
        if (e_ != cudaSuccess) {                                       \
            fail("d3q19_create: %s -> %s", #call, cudaGetErrorString(e_)); \
            d3q19_destroy(h);                                          \
            return 1;                                                  \
        }                                                              \
    } while (0)
    int lo = 0, hi = 0;
    CKH(cudaDeviceGetStreamPriorityRange(&lo, &hi));
    CKH(cudaStreamCreateWithPriority(&h->sc, cudaStreamNonBlocking, lo));
    CKH(cudaStreamCreateWithPriority(&h->sx, cudaStreamNonBlocking, hi));
    CKH(cudaEventCreateWithFlags(&h->evB, cudaEventDisableTiming));
    CKH(cudaEventCreateWithFlags(&h->evX, cudaEventDisableTiming));
    for (int i = 0; i < 2; ++i) {
        CKH(cudaEventCreateWithFlags(&h->evC[i], cudaEventDisableTiming));
        CKH(cudaEventCreateWithFlags(&h->evS[i], cudaEventDisableTiming));
    }
    CKH(cudaEventCreate(&h->t0));
    CKH(cudaEventCreate(&h->t1));
    const size_t PAD = POP_PAD;
    const size_t fbytes = ((size_t)NPOP * g.slab + 2 * PAD) * sizeof(double);
    h->idx32 = (unsigned long long)g.slab + 2 < 0xffffffffull && !cfg->force_idx64;
    if (cfg->halo_timeout_s > 0) h->halo_timeout_ns = (unsigned long long)cfg->halo_timeout_s * 1000000000ull;
    if (cfg->halo_split_min > 0) h->halo_split_min = cfg->halo_split_min;
    CKH(cudaMalloc(&h->A_alloc, fbytes));
    CKH(cudaMemsetAsync(h->A_alloc, 0, fbytes, h->sc));
    h->A = h->A_alloc + PAD;
    if (h->cfg.scheme == D3Q19_SCHEME_AB) {
        CKH(cudaMalloc(&h->B_alloc, fbytes));
        CKH(cudaMemsetAsync(h->B_alloc, 0, fbytes, h->sc));
        h->B = h->B_alloc + PAD;
    }
    CKH(cudaMalloc(&h->scal, 64 * sizeof(double)));
    CKH(cudaMemsetAsync(h->scal, 0, 64 * sizeof(double), h->sc));
    if (cfg->nranks > 1) {
        NcclApi &n = nccl_api();
        if (const char *m = n.load()) { fail("d3q19_create: %s", m); d3q19_destroy(h); return 1; }
        NcclUniqueId id;
        memcpy(id.internal, cfg->nccl_id, 128);
        // d3q19_config.nccl_max_ctas > 0 caps the CTAs of NCCL's send/recv kernel.  Measured (2 B200, 32-plane slabs,
        // profiles/r02e_two_gpus.md): every cap is a loss -- 1-2 CTAs 0.655 ms per step, 4: 0.419, 8: 0.298, 32 = no cap: 0.230 --
        // the exchange becomes too slow to hide long before it stops disturbing the interior kernel.  Default: NCCL's own choice.
        NcclConfig218 nc = nccl_config_max_ctas(cfg->nccl_max_ctas);
        int e = (cfg->nccl_max_ctas > 0 && n.CommInitRankConfig) ? n.CommInitRankConfig(&h->comm, cfg->nranks, id, cfg->rank, &nc)
                                                                 : n.CommInitRank(&h->comm, cfg->nranks, id, cfg->rank);
        if (e != 0) { fail("d3q19_create: ncclCommInitRank -> %s", n.GetErrorString(e)); h->comm = nullptr; d3q19_destroy(h); return 1; }
        const size_t fb = (size_t)5 * g.plane * sizeof(double);
        CKH(cudaMalloc(&h->send_up, fb)); CKH(cudaMalloc(&h->send_dn, fb));
        CKH(cudaMalloc(&h->recv_lo, fb)); CKH(cudaMalloc(&h->recv_hi, fb));
    }
    CKH(cudaStreamSynchronize(h->sc));
#undef CKH
    *out = h;
    return 0;
}

extern "C" int d3q19_sync(d3q19_handle *h) {
    CK(cudaSetDevice(h->cfg.device));
    if (h->halo_on) RK_(wait_exchange(h));       // the neighbours' stores of the last step have landed
    CK(cudaStreamSynchronize(h->sc));
    CK(cudaStreamSynchronize(h->sx));
    if (h->halo_on) {
        // watchdog of the flag waits (kernels.cuh halo_spin): a neighbour's flag that never came
        unsigned int bad = 0;
        CK(cudaMemcpy(&bad, h->halo_flags + 8, sizeof bad, cudaMemcpyDeviceToHost));
        if (bad)
            return fail("d3q19_sync: rank %d waited more than %.0f s for a neighbour's halo flag of step %u; the populations "
                        "are void (a neighbour rank died or peer memory is broken)", h->cfg.rank,
                        (double)h->halo_timeout_ns * 1e-9, bad);
    }
    return check_link_overflow(h, "d3q19_sync");
}

// ---- halo in peer memory: cudaIpc bootstrap ---------------------------------------------------------------
// blob layout (D3Q19_IPC_BYTES = 256): [0,64) handle of the first population array, [64,128) of the
// second (AB) or zeros, [128,192) of the flag words, then int32 lz, int32 scheme, int64 slab.
extern "C" int d3q19_ipc_export(d3q19_handle *h, unsigned char *blob) {
    CK(cudaSetDevice(h->cfg.device));
    if (h->cfg.nranks < 2) return fail("d3q19_ipc_export: needs nranks > 1");
    if (!h->halo_flags) {
        CK(cudaMalloc(&h->halo_flags, 64 * sizeof(unsigned int)));
        CK(cudaMemset(h->halo_flags, 0, 64 * sizeof(unsigned int)));
    }
    memset(blob, 0, D3Q19_IPC_BYTES);
    cudaIpcMemHandle_t mh;
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "cudaIpcMemHandle_t is 64 bytes");
    CK(cudaIpcGetMemHandle(&mh, h->A_alloc)); memcpy(blob, &mh, 64);
    if (h->B_alloc) { CK(cudaIpcGetMemHandle(&mh, h->B_alloc)); memcpy(blob + 64, &mh, 64); }
    CK(cudaIpcGetMemHandle(&mh, h->halo_flags)); memcpy(blob + 128, &mh, 64);
    int32_t meta[2] = {h->g.lz, h->cfg.scheme};
    memcpy(blob + 192, meta, sizeof meta);
    long long slab = h->g.slab;
    memcpy(blob + 200, &slab, sizeof slab);
    // A and B may have been swapped by earlier steps: say which allocation is the current source
    int32_t a_is_first = (h->A == h->A_alloc + POP_PAD) ? 1 : 0;
    memcpy(blob + 208, &a_is_first, sizeof a_is_first);
    return 0;
}

extern "C" int d3q19_ipc_connect(d3q19_handle *h, const unsigned char *blobs) {
    CK(cudaSetDevice(h->cfg.device));
    if (h->cfg.nranks < 2) return fail("d3q19_ipc_connect: needs nranks > 1");
    if (h->cfg.ipart && h->halo_mode != D3Q19_HALO_PUT)
        return fail("d3q19_ipc_connect: with particles (ipart) the faces travel by NCCL or by the copy engines -- call "
                    "d3q19_set_halo_mode(h, D3Q19_HALO_PUT) first");
    if (h->g.lz < 2) return fail("d3q19_ipc_connect: slabs must be at least 2 planes thick");
    if (!h->halo_flags) return fail("d3q19_ipc_connect: call d3q19_ipc_export first");
    if (h->halo_on) return fail("d3q19_ipc_connect: already connected");
    RK_(wait_exchange(h));
    const int nb[2] = {(h->cfg.rank + h->cfg.nranks - 1) % h->cfg.nranks, (h->cfg.rank + 1) % h->cfg.nranks};
    // Opening can fail on one rank only (no peer access between two devices, a foreign container ...):
    // every rank tries, then all agree through an all-reduce, so that either all use the peer halo or none.
    int failed = 0;
    std::string why;
    for (int d = 0; d < 2 && !failed; ++d) {
        const unsigned char *b = blobs + (size_t)nb[d] * D3Q19_IPC_BYTES;
        int32_t meta[2]; long long slab; int32_t a_first;
        memcpy(meta, b + 192, sizeof meta); memcpy(&slab, b + 200, sizeof slab); memcpy(&a_first, b + 208, sizeof a_first);
        if (meta[1] != h->cfg.scheme) { failed = 1; why = "neighbour runs another storage scheme"; break; }
        h->peer_lz[d] = meta[0];
        h->peer_slab[d] = slab;
        if (d == 1 && nb[1] == nb[0]) {                 // two ranks: both neighbours are the same process
            for (int a = 0; a < 3; ++a) h->peer_base[1][a] = h->peer_base[0][a];
        } else {
            cudaIpcMemHandle_t mh;
            const unsigned char zero[64] = {0};
            for (int a = 0; a < 3; ++a) {
                if (a == 1 && !memcmp(b + 64, zero, 64)) continue;
                memcpy(&mh, b + 64 * a, 64);
                cudaError_t e = cudaIpcOpenMemHandle(&h->peer_base[d][a], mh, cudaIpcMemLazyEnablePeerAccess);
                if (e != cudaSuccess) {
                    failed = 1; why = std::string("cudaIpcOpenMemHandle -> ") + cudaGetErrorString(e);
                    h->peer_base[d][a] = nullptr;
                    cudaGetLastError();
                    break;
                }
            }
        }
        if (failed) break;
        double *first = (double *)h->peer_base[d][0] + POP_PAD;
        double *second = h->peer_base[d][1] ? (double *)h->peer_base[d][1] + POP_PAD : nullptr;
        h->peer_A[d] = a_first ? first : second;
        h->peer_B[d] = a_first ? second : first;
        h->peer_flags[d] = (unsigned int *)h->peer_base[d][2];
    }
    // agreement + barrier: everybody has opened everybody before the first remote store
    double nfail = (double)failed;
    CK(cudaMemcpyAsync(h->scal + 32, &nfail, sizeof nfail, cudaMemcpyHostToDevice, h->sc));
    NK(nccl_api().AllReduce(h->scal + 32, h->scal + 32, 1, NCCL_FLOAT64, NCCL_SUM, h->comm, h->sc));
    h->n_nccl++;
    CK(cudaMemcpyAsync(&nfail, h->scal + 32, sizeof nfail, cudaMemcpyDeviceToHost, h->sc));
    CK(cudaStreamSynchronize(h->sc));
    if (nfail > 0.0) {
        for (int d = 0; d < 2; ++d)
            for (int a = 0; a < 3; ++a) {
                if (h->peer_base[d][a] && !(d == 1 && h->peer_base[0][a] == h->peer_base[1][a])) cudaIpcCloseMemHandle(h->peer_base[d][a]);
            }
        for (int d = 0; d < 2; ++d)
            for (int a = 0; a < 3; ++a) h->peer_base[d][a] = nullptr;
        fail("d3q19_ipc_connect: peer memory unavailable on %d rank(s)%s%s; the halo stays on NCCL", (int)nfail,
             failed ? ": " : "", failed ? why.c_str() : "");
        return 2;
    }
    h->halo_on = true;
    return 0;
}

extern "C" int d3q19_set_halo_mode(d3q19_handle *h, int32_t mode) {
    if (mode != D3Q19_HALO_FUSED && mode != D3Q19_HALO_PUT) return fail("d3q19_set_halo_mode: unknown mode %d", mode);
    if (h->halo_on && h->halo_epoch > 0) return fail("d3q19_set_halo_mode: steps were already taken in the other mode");
    h->halo_mode = mode;
    return 0;
}

// collective operations that rewrite the populations end with a barrier in halo mode, so that no
// neighbour starts storing into planes that are still being filled
static int halo_barrier(d3q19_handle *h) {
    if (!h->halo_on || !h->comm) return 0;
    NK(nccl_api().AllReduce(h->scal + 32, h->scal + 32, 1, NCCL_FLOAT64, NCCL_SUM, h->comm, h->sc));
    h->n_nccl++;
    CK(cudaStreamSynchronize(h->sc));
    return 0;
}

// ---- state transfer ----------------------------------------------------------------------------------
extern "C" int d3q19_upload_f(d3q19_handle *h, const double *f_aos) {
    CK(cudaSetDevice(h->cfg.device));
    RK_(ensure_stage(h));
    RK_(wait_exchange(h));
    const Geom &g = h->g;
    const size_t per_plane = (size_t)NPOP * g.lx * g.ly;
    // canonical populations land in A (AA) or in B (AB, un-streamed into A below)
    double *dst = h->cfg.scheme == D3Q19_SCHEME_AB ? h->B : h->A;
    int c = 0;
    for (int z0 = 0; z0 < g.lz; z0 += h->stage_planes, ++c) {
        const int nz = (z0 + h->stage_planes <= g.lz) ? h->stage_planes : g.lz - z0;
        const int b = c & 1;
        if (c >= 2) CK(cudaStreamWaitEvent(h->sx, h->evS[b], 0));          // stage b free again
        CK(cudaMemcpyAsync(h->stage[b], f_aos + (size_t)z0 * per_plane, (size_t)nz * per_plane * sizeof(double),
                           cudaMemcpyHostToDevice, h->sx));
        CK(cudaEventRecord(h->evC[b], h->sx));
        CK(cudaStreamWaitEvent(h->sc, h->evC[b], 0));
        k_scatter_aos<<<grid_nodes(h, nz), BLOCK_X, 0, h->sc>>>(g, dst, h->stage[b], 1 + z0);
        CK(cudaGetLastError());
        CK(cudaEventRecord(h->evS[b], h->sc));
        h->n_other_kernels++;
    }
    if (h->cfg.scheme == D3Q19_SCHEME_AB) {
        if (h->cfg.nranks > 1) {
            // canonical neighbours of the boundary planes: plane lz (c_z=-1 slots) -> +z ghost 0,
            // plane 1 (c_z=+1 slots) -> -z ghost lz+1
            CK(cudaEventRecord(h->evB, h->sc));
            CK(cudaStreamWaitEvent(h->sx, h->evB, 0));
            RK_(exchange_faces(h, h->B, g.lz, SLOTS_MZ, 0, 1, SLOTS_PZ, g.lz + 1, 0, h->sx));
            CK(cudaEventRecord(h->evX, h->sx));
            CK(cudaStreamWaitEvent(h->sc, h->evX, 0));
        }
        k_unstream<<<grid_nodes(h, g.lz), BLOCK_X, 0, h->sc>>>(g, h->B, h->A);
        CK(cudaGetLastError());
        h->n_other_kernels++;
        if (h->cfg.nranks > 1) {
            CK(cudaEventRecord(h->evB, h->sc));
            CK(cudaStreamWaitEvent(h->sx, h->evB, 0));
            RK_(exchange_after_step(h, STEP_AB, h->A, h->sx));
            CK(cudaEventRecord(h->evX, h->sx));
            h->exchange_pending = true;
        }
    }
    h->phase = 0;
    // the shim's coherence flags live in the primitives, so that raw and shim calls can be mixed
    h->shim.f_dev_valid = true;
    h->shim.f_host_valid = h->shim.bound && f_aos == h->shim.a.f;
    h->shim.macro_dev_valid = false;
    CK(cudaStreamSynchronize(h->sx));
    CK(cudaStreamSynchronize(h->sc));
    return halo_barrier(h);
}

template <int RKIND>
static int download_f_impl(d3q19_handle *h, double *f_aos) {
    const Geom &g = h->g;
    const size_t per_plane = (size_t)NPOP * g.lx * g.ly;
    int c = 0;
    for (int z0 = 0; z0 < g.lz; z0 += h->stage_planes, ++c) {
        const int nz = (z0 + h->stage_planes <= g.lz) ? h->stage_planes : g.lz - z0;
        const int b = c & 1;
        if (c >= 2) CK(cudaStreamWaitEvent(h->sc, h->evC[b], 0));          // copy-out of stage b finished
        k_gather_aos<RKIND><<<grid_nodes(h, nz), BLOCK_X, 0, h->sc>>>(g, h->A, h->stage[b], 1 + z0);
        CK(cudaGetLastError());
        CK(cudaEventRecord(h->evS[b], h->sc));
        CK(cudaStreamWaitEvent(h->sx, h->evS[b], 0));
        CK(cudaMemcpyAsync(f_aos + (size_t)z0 * per_plane, h->stage[b], (size_t)nz * per_plane * sizeof(double),
                           cudaMemcpyDeviceToHost, h->sx));
        CK(cudaEventRecord(h->evC[b], h->sx));
        h->n_other_kernels++;
    }
    CK(cudaStreamSynchronize(h->sx));
    CK(cudaStreamSynchronize(h->sc));
    return 0;
}

extern "C" int d3q19_download_f(d3q19_handle *h, double *f_aos) {
    CK(cudaSetDevice(h->cfg.device));
    RK_(ensure_stage(h));
    RK_(wait_exchange(h));
    int rc;
    switch (read_kind(h)) {
    case READ_DIRECT: rc = download_f_impl<READ_DIRECT>(h, f_aos); break;
    case READ_PULL_NAT: rc = download_f_impl<READ_PULL_NAT>(h, f_aos); break;
    default: rc = download_f_impl<READ_PULL_SWAP>(h, f_aos); break;
    }
    if (rc == 0 && h->shim.bound && f_aos == h->shim.a.f) h->shim.f_host_valid = true;
    return rc;
}

// pitched device field <-> host (lx,ly,lz) through the staging buffer
static int field_to_host(d3q19_handle *h, const double *dev, double *host) {
    if (!host) return 0;
    const Geom &g = h->g;
    RK_(ensure_stage(h));
    if (g.xp == g.lx) {
        CK(cudaMemcpyAsync(host, dev, h->nfield * sizeof(double), cudaMemcpyDeviceToHost, h->sc));
        return 0;
    }
    CK(cudaMemcpy2DAsync(host, (size_t)g.lx * sizeof(double), dev, (size_t)g.xp * sizeof(double),
                         (size_t)g.lx * sizeof(double), (size_t)g.ly * g.lz, cudaMemcpyDeviceToHost, h->sc));
    return 0;
}
static int field_from_host(d3q19_handle *h, double *dev, const double *host) {
    if (!host) return 0;
    const Geom &g = h->g;
    CK(cudaMemcpy2DAsync(dev, (size_t)g.xp * sizeof(double), host, (size_t)g.lx * sizeof(double),
                         (size_t)g.lx * sizeof(double), (size_t)g.ly * g.lz, cudaMemcpyHostToDevice, h->sc));
    return 0;
}

extern "C" int d3q19_set_macro(d3q19_handle *h, const double *rho, const double *ux, const double *uy, const double *uz) {
    CK(cudaSetDevice(h->cfg.device));
    RK_(ensure_macro_arrays(h));
    RK_(field_from_host(h, h->rho, rho)); RK_(field_from_host(h, h->ux, ux));
    RK_(field_from_host(h, h->uy, uy)); RK_(field_from_host(h, h->uz, uz));
    CK(cudaStreamSynchronize(h->sc));
    h->shim.macro_dev_valid = false;    // the device arrays are what the caller set, not the moments of f
    return 0;
}

extern "C" int d3q19_download_macro(d3q19_handle *h, double *rho, double *ux, double *uy, double *uz) {
    CK(cudaSetDevice(h->cfg.device));
    RK_(ensure_macro_arrays(h));
    RK_(field_to_host(h, h->rho, rho)); RK_(field_to_host(h, h->ux, ux));
    RK_(field_to_host(h, h->uy, uy)); RK_(field_to_host(h, h->uz, uz));
    CK(cudaStreamSynchronize(h->sc));
    return 0;
}

// ---- vortcalc (saveload.f90:3929-4054) ----------------------------------------------------------------
// Works on the device rho,u arrays d3q19_macrovar made (the reference's vortcalc reads the ux,uy,uz macrovar
// left behind).  The z phase of exchng8 (:4039-4045) is three contiguous planes each way over NCCL; its y
// phase is the periodic index wrap.
extern "C" int d3q19_vortcalc(d3q19_handle *h) {
    CK(cudaSetDevice(h->cfg.device));
    if (!h->rho) return fail("d3q19_vortcalc: no velocity field on the device -- call d3q19_macrovar first");
    if (h->g.ly < 2 || h->g.lz < 2) return fail("d3q19_vortcalc: needs ly >= 2 and lz >= 2 (saveload.f90:3982-3999)");
    const Geom &g = h->g;
    if (!h->vort) CK(cudaMalloc(&h->vort, 3 * h->nfield * sizeof(double)));
    VortParams p;
    memset(&p, 0, sizeof p);
    p.g = g;
    p.ux = h->ux; p.uy = h->uy; p.uz = h->uz;
    p.ox = h->vort; p.oy = h->vort + h->nfield; p.oz = h->vort + 2 * h->nfield;
    p.solid = h->solid; p.isnodes = h->isn; p.omgp = h->omgp;
    if (h->solid && (!h->isn || !h->omgp)) return fail("d3q19_vortcalc: solid nodes need isnodes and the particle table");
    if (h->cfg.nranks > 1) {
        if (!h->vort_halo) CK(cudaMalloc(&h->vort_halo, 6 * (size_t)g.plane * sizeof(double)));
        const int up = (h->cfg.rank + 1) % h->cfg.nranks, dn = (h->cfg.rank + h->cfg.nranks - 1) % h->cfg.nranks;
        const size_t cnt = (size_t)g.plane;
        const double *fld[3] = {h->ux, h->uy, h->uz};
        NcclApi &n = nccl_api();
        NK(n.GroupStart());
        for (int c = 0; c < 3; ++c) {
            NK(n.Send(fld[c] + (size_t)(g.lz - 1) * g.plane, cnt, NCCL_FLOAT64, up, h->comm, h->sc));   // my plane lz  -> mzp's tmpu?F
            NK(n.Send(fld[c], cnt, NCCL_FLOAT64, dn, h->comm, h->sc));                                   // my plane 1   -> mzm's tmpu?B
            NK(n.Recv(h->vort_halo + (size_t)c * g.plane, cnt, NCCL_FLOAT64, dn, h->comm, h->sc));
            NK(n.Recv(h->vort_halo + (size_t)(3 + c) * g.plane, cnt, NCCL_FLOAT64, up, h->comm, h->sc));
        }
        NK(n.GroupEnd());
        h->n_nccl += 12;
        p.zlo = h->vort_halo; p.zhi = h->vort_halo + 3 * (size_t)g.plane;
    }
    k_vortcalc<<<grid_nodes(h, g.lz), BLOCK_X, 0, h->sc>>>(p);
    CK(cudaGetLastError());
    h->n_other_kernels++;
    return 0;
}

extern "C" int d3q19_download_vort(d3q19_handle *h, double *ox, double *oy, double *oz) {
    CK(cudaSetDevice(h->cfg.device));
    if (!h->vort) return fail("d3q19_download_vort: call d3q19_vortcalc first");
    RK_(field_to_host(h, h->vort, ox)); RK_(field_to_host(h, h->vort + h->nfield, oy));
    RK_(field_to_host(h, h->vort + 2 * h->nfield, oz));
    CK(cudaStreamSynchronize(h->sc));
    return 0;
}

// ---- local strain rate (sijstat00's first loop nest, saveload.f90:2031-2091) ---------------------------------------
// Like vortcalc it works on the rho,u arrays d3q19_macrovar made (the reference reads the arrays macrovar left behind)
// and on the current populations; node-local, so no exchange.
extern "C" int d3q19_sijstat(d3q19_handle *h) {
    CK(cudaSetDevice(h->cfg.device));
    if (!h->rho) return fail("d3q19_sijstat: no rho,u on the device -- call d3q19_macrovar first");
    RK_(ensure_mask(h));
    RK_(wait_exchange(h));
    if (!h->sij2) CK(cudaMalloc(&h->sij2, h->nfield * sizeof(double)));
    SijParams p;
    memset(&p, 0, sizeof p);
    p.g = h->g; p.A = h->A;
    p.rho = h->rho; p.ux = h->ux; p.uy = h->uy; p.uz = h->uz;
    p.solid = h->solid;
    p.s1 = h->cfg.s1; p.s9 = h->cfg.s9;
    p.sij2 = h->sij2;
    const dim3 gr = grid_nodes(h, h->g.lz);
    switch (read_kind(h)) {
    case READ_DIRECT: k_sijstat<READ_DIRECT><<<gr, BLOCK_X, 0, h->sc>>>(p); break;
    case READ_PULL_NAT: k_sijstat<READ_PULL_NAT><<<gr, BLOCK_X, 0, h->sc>>>(p); break;
    default: k_sijstat<READ_PULL_SWAP><<<gr, BLOCK_X, 0, h->sc>>>(p); break;
    }
    CK(cudaGetLastError());
    h->n_other_kernels++;
    return 0;
}

extern "C" int d3q19_download_sij2(d3q19_handle *h, double *sij2) {
    CK(cudaSetDevice(h->cfg.device));
    if (!h->sij2) return fail("d3q19_download_sij2: call d3q19_sijstat first");
    RK_(field_to_host(h, h->sij2, sij2));
    CK(cudaStreamSynchronize(h->sc));
    return 0;
}

// ---- forcing ---------------------------------------------------------------------------------------
extern "C" int d3q19_set_force_uniform(d3q19_handle *h, double fx, double fy, double fz) {
    CK(cudaSetDevice(h->cfg.device));
    h->Fx = fx; h->Fy = fy; h->Fz = fz;
    if (h->ffx) {
        CK(cudaStreamSynchronize(h->sc));
        cudaFree(h->ffx); cudaFree(h->ffy); cudaFree(h->ffz);
        h->ffx = h->ffy = h->ffz = nullptr;
    }
    return 0;
}

extern "C" int d3q19_set_force_field(d3q19_handle *h, const double *fx, const double *fy, const double *fz) {
    CK(cudaSetDevice(h->cfg.device));
    if (!fx || !fy || !fz) return fail("d3q19_set_force_field: null array");
    const size_t bytes = h->nfield * sizeof(double);
    if (!h->ffx) {
        CK(cudaMalloc(&h->ffx, bytes)); CK(cudaMalloc(&h->ffy, bytes)); CK(cudaMalloc(&h->ffz, bytes));
        CK(cudaMemsetAsync(h->ffx, 0, bytes, h->sc)); CK(cudaMemsetAsync(h->ffy, 0, bytes, h->sc));
        CK(cudaMemsetAsync(h->ffz, 0, bytes, h->sc));
    }
    RK_(field_from_host(h, h->ffx, fx)); RK_(field_from_host(h, h->ffy, fy)); RK_(field_from_host(h, h->ffz, fz));
    CK(cudaStreamSynchronize(h->sc));
    return 0;
}

// FORCINGP (collision.f90:529-602) on the device: fills the force field for step `istep` without touching
// the host (the reference's host loop + an upload would move 24 B per node over PCIe every step).
extern "C" int d3q19_forcingp(d3q19_handle *h, int32_t istep, double force_in_y) {
    CK(cudaSetDevice(h->cfg.device));
    const size_t bytes = h->nfield * sizeof(double);
    if (!h->ffx) {
        CK(cudaMalloc(&h->ffx, bytes)); CK(cudaMalloc(&h->ffy, bytes)); CK(cudaMalloc(&h->ffz, bytes));
    }
    ForcingpParams p;
    memset(&p, 0, sizeof p);
    const Geom &g = h->g;
    p.lx = g.lx; p.ly = g.ly; p.lz = g.lz; p.xp = g.xp;
    p.nx = h->cfg.nx; p.ny = h->cfg.ny; p.nz = h->cfg.nz; p.globalz = h->cfg.globalz;
    const double pi = 4.0 * atan(1.0);                          // var_inc.f90:71
    p.pi2 = 2.0 * pi;
    const double Tpd = 2000.0;                                  // :538
    p.beta9 = 3.0; p.gamma9 = 2.0; p.phase9 = 0.25; p.ixs0 = 2; // :540-545
    p.ihh = (g.lx / 2) / 2;                                     // lxh/2, var_inc.f90:53
    p.force_in_y = force_in_y;
    p.Amp0 = 40.00 * p.beta9 / (double)p.ny * sin(p.pi2 * (double)istep / Tpd);   // :543
    if (p.ihh < 1) return fail("d3q19_forcingp: channel too narrow (lx = %d)", g.lx);
    p.fx = h->ffx; p.fy = h->ffy; p.fz = h->ffz;
    k_forcingp<<<grid_nodes(h, g.lz), BLOCK_X, 0, h->sc>>>(p);
    CK(cudaGetLastError());
    h->n_other_kernels++;
    return 0;
}

// device force field -> host arrays in the reference's layout (force_realx/y/z(lx,ly,lz), para.f90:443-445)
extern "C" int d3q19_download_force_field(d3q19_handle *h, double *fx, double *fy, double *fz) {
    CK(cudaSetDevice(h->cfg.device));
    if (!h->ffx) return fail("d3q19_download_force_field: no force field is set (uniform force)");
    RK_(field_to_host(h, h->ffx, fx)); RK_(field_to_host(h, h->ffy, fy)); RK_(field_to_host(h, h->ffz, fz));
    CK(cudaStreamSynchronize(h->sc));
    return 0;
}

// ---- the step ----------------------------------------------------------------------------------------
template <int SK, bool STRICT, bool GENERIC>
static int launch_step_range(d3q19_handle *h, const StepParams &p0, int z0, int nplanes, cudaStream_t s, int zstride = 1) {
    if (nplanes <= 0) return 0;
    StepParams p = p0;
    p.z0 = z0;
    p.zstride = zstride;
    if (h->idx32) k_step<SK, STRICT, GENERIC, uint32_t><<<grid_nodes(h, nplanes), BLOCK_X, 0, s>>>(p);
    else k_step<SK, STRICT, GENERIC, unsigned long long><<<grid_nodes(h, nplanes), BLOCK_X, 0, s>>>(p);
    CK(cudaGetLastError());
    h->n_step_kernels++;
    return 0;
}

// Halo stored into the neighbours' memory.  Thin slabs: ONE launch for the whole slab, boundary planes first
// in block order.  Thick slabs (lz >= halo_split_min): the two boundary planes run the halo instantiation,
// the interior the plain one (which needs fewer registers -- the AA odd step keeps its 4 CTAs/SM); the
// order of the two launches is free, no node of one reads or writes an address the other writes.
template <int SK, bool STRICT, bool GENERIC>
static int launch_step_halo(d3q19_handle *h, const StepParams &p0) {
    StepParams p = p0;
    Halo &q = p.halo;
    const bool ab = SK == STEP_AB;
    // AB: this step writes B here and in the neighbours; AA: the single array
    q.peer_dn = ab ? h->peer_B[0] : h->peer_A[0];
    q.peer_up = ab ? h->peer_B[1] : h->peer_A[1];
    q.slab_dn = h->peer_slab[0]; q.slab_up = h->peer_slab[1];
    q.lz_dn = h->peer_lz[0];
    q.wait_lo = h->halo_flags; q.wait_hi = h->halo_flags + 1;
    q.sig_dn = h->peer_flags[0] + 1;      // the lower neighbour's wait_hi
    q.sig_up = h->peer_flags[1];          // the upper neighbour's wait_lo
    q.ctr = h->halo_flags + 2;
    q.err = h->halo_flags + 8;
    q.timeout_ns = h->halo_timeout_ns;
    q.epoch = ++h->halo_epoch;
    const bool split = h->g.lz >= h->halo_split_min && h->g.lz > 2;
    const dim3 gr = grid_nodes(h, split ? 2 : h->g.lz);
    q.nblk_face = gr.x * gr.y;
    if (h->idx32) k_step<SK, STRICT, GENERIC, uint32_t, true><<<gr, BLOCK_X, 0, h->sc>>>(p);
    else k_step<SK, STRICT, GENERIC, unsigned long long, true><<<gr, BLOCK_X, 0, h->sc>>>(p);
    CK(cudaGetLastError());
    h->n_step_kernels++;
    if (split) RK_((launch_step_range<SK, STRICT, GENERIC>(h, p0, 2, h->g.lz - 2, h->sc)));
    if (ab) {                             // the neighbours swap their arrays in lockstep
        for (int d = 0; d < 2; ++d) { double *t = h->peer_A[d]; h->peer_A[d] = h->peer_B[d]; h->peer_B[d] = t; }
    }
    return 0;
}

// Copy-engine transport (D3Q19_HALO_PUT, the default of bench.py): plain step kernels, boundary planes first; the copy
// engines move both faces into the neighbours' arrays on the second stream and a one-thread kernel raises their flags.
// Ordering: this step's boundary kernel runs after (a) my own previous copies have finished reading my planes (evX) and
// (b) both neighbours' previous copies have landed (flags >= epoch-1); (b) also implies the neighbours are done reading
// the ghost planes this step's copies overwrite.
template <int SK, bool STRICT, bool GENERIC>
static int launch_step_put(d3q19_handle *h, const StepParams &p, double *written) {
    const int lz = h->g.lz;
    const Geom &g = h->g;
    const unsigned int epoch = ++h->halo_epoch;
    if (h->put_pending) { CK(cudaStreamWaitEvent(h->sc, h->evX, 0)); h->put_pending = false; }
    // (b) is waited for by the blocks of the boundary launch themselves (k_step: "the planes next to a face"); a slab of one
    // or two planes is all boundary
    StepParams pb = p;
    if (epoch > 1) {
        pb.halo.wait_lo = h->halo_flags; pb.halo.wait_hi = h->halo_flags + 1;
        pb.halo.epoch = epoch; pb.halo.err = h->halo_flags + 8; pb.halo.timeout_ns = h->halo_timeout_ns;
    }
    trace_mark(h, 0, h->sc);
    // (One launch for the whole slab with the boundary planes first in block order and a device-side "planes done" signal
    //  for the copy stream was measured as well: 0.216 ms per step against 0.206 ms for the two launches below on 32-plane
    //  slabs -- the instantiation that counts blocks is slower than the plain one.  Removed.  profiles/r02e_two_gpus.md)
    if (lz > 2) {
        RK_((launch_step_range<SK, STRICT, GENERIC>(h, pb, 1, 2, h->sc, lz - 1)));  // planes 1 and lz in one launch
        CK(cudaEventRecord(h->evB, h->sc));
        trace_mark(h, 1, h->sc);
        RK_((launch_step_range<SK, STRICT, GENERIC>(h, p, 2, lz - 2, h->sc)));
    } else {
        RK_((launch_step_range<SK, STRICT, GENERIC>(h, pb, 1, lz, h->sc)));
        CK(cudaEventRecord(h->evB, h->sc));
        trace_mark(h, 1, h->sc);
    }
    trace_mark(h, 2, h->sc);
    CK(cudaStreamWaitEvent(h->sx, h->evB, 0));
    // The faces travel by the COPY ENGINES: one population of one z plane is plane = xp*ly contiguous doubles, so each
    // of the five crossing populations goes from where it lies in my array to where it belongs in the neighbour's
    // (cudaIpc-mapped) array with one device-to-device copy over NVLink -- no SM takes part, nothing competes with the
    // interior kernel for issue slots or registers (measured: NCCL's send/recv CTAs and a copy kernel both slow the
    // concurrent interior launch by 35-50 us on 32-plane slabs, profiles/r02b_timeline_thin.jsonl).  After an in-place
    // odd step the ghost planes go back into the neighbour's REAL plane and must leave the wall-adjacent nodes of the
    // c_x = +-1 populations alone (collision.f90:361-368): those two populations per face are pitched 2-D copies of
    // lx-1 columns.  Copies of one stream complete in order; a one-thread kernel then raises the neighbours' flags.
    const bool ab = SK == STEP_AB;
    double *dstA[2] = {ab ? h->peer_B[1] : h->peer_A[1], ab ? h->peer_B[0] : h->peer_A[0]};     // upper, lower neighbour
    const long long slab_dst[2] = {h->peer_slab[1], h->peer_slab[0]};
    int zsrc[2], zdst[2], excl = 0;
    FaceSlots sl[2];
    switch (SK) {                                       // the planes and slots of exchange_after_step
    case STEP_AB:      zsrc[0] = lz; sl[0] = SLOTS_PZ; zdst[0] = 0; zsrc[1] = 1; sl[1] = SLOTS_MZ; zdst[1] = h->peer_lz[0] + 1; break;
    case STEP_AA_EVEN: zsrc[0] = lz; sl[0] = SLOTS_MZ; zdst[0] = 0; zsrc[1] = 1; sl[1] = SLOTS_PZ; zdst[1] = h->peer_lz[0] + 1; break;
    default:           zsrc[0] = lz + 1; sl[0] = SLOTS_PZ; zdst[0] = 1; zsrc[1] = 0; sl[1] = SLOTS_MZ; zdst[1] = h->peer_lz[0];
                       excl = 1; break;
    }
    const size_t pl = (size_t)g.plane;
    for (int q = 0; q < 5; ++q) {
        for (int d = 0; d < 2; ++d) {
            const int slot = sl[d].s[q];
            const double *src = written + (size_t)slot * g.slab + (size_t)zsrc[d] * pl;
            double *dst = dstA[d] + (size_t)slot * slab_dst[d] + (size_t)zdst[d] * pl;
            const int cx = d3q::dir_cx_rt(slot);
            if (excl && cx != 0) {
                const size_t off = cx > 0 ? 1 : 0;      // c_x = +1 skips x = 0, c_x = -1 skips x = lx-1
                CK(cudaMemcpy2DAsync(dst + off, (size_t)g.xp * sizeof(double), src + off, (size_t)g.xp * sizeof(double),
                                     (size_t)(g.lx - 1) * sizeof(double), (size_t)g.ly, cudaMemcpyDeviceToDevice, h->sx));
            } else {
                CK(cudaMemcpyAsync(dst, src, pl * sizeof(double), cudaMemcpyDeviceToDevice, h->sx));
            }
        }
    }
    k_flag_raise<<<1, 1, 0, h->sx>>>(h->peer_flags[1], h->peer_flags[0] + 1, epoch);    // upper's wait_lo, lower's wait_hi
    CK(cudaGetLastError());
    CK(cudaEventRecord(h->evX, h->sx));
    trace_mark(h, 3, h->sx);
    trace_next(h);
    h->put_pending = true;
    h->n_other_kernels += 1;
    h->n_copies += 10;
    if (ab) {                             // the neighbours swap their arrays in lockstep
        for (int d = 0; d < 2; ++d) { double *t = h->peer_A[d]; h->peer_A[d] = h->peer_B[d]; h->peer_B[d] = t; }
    }
    return 0;
}

template <int SK, bool STRICT, bool GENERIC>
static int step_impl(d3q19_handle *h, const StepParams &p, double *written) {
    const int lz = h->g.lz;
    if (h->cfg.nranks == 1) {
        trace_mark(h, 0, h->sc); trace_mark(h, 1, h->sc);
        RK_((launch_step_range<SK, STRICT, GENERIC>(h, p, 1, lz, h->sc)));
        trace_mark(h, 2, h->sc); trace_mark(h, 3, h->sc);
        trace_next(h);
        return 0;
    }
    if (h->halo_on) {
        if (h->exchange_pending) {          // a ghost fill by NCCL (upload, init) precedes the first halo step
            CK(cudaStreamWaitEvent(h->sc, h->evX, 0));
            h->exchange_pending = false;
        }
        if (h->halo_mode == D3Q19_HALO_PUT) return launch_step_put<SK, STRICT, GENERIC>(h, p, written);
        trace_mark(h, 0, h->sc); trace_mark(h, 1, h->sc);
        RK_((launch_step_halo<SK, STRICT, GENERIC>(h, p)));
        trace_mark(h, 2, h->sc); trace_mark(h, 3, h->sc);
        trace_next(h);
        return 0;
    }
    // boundary planes first, so that their faces travel while the interior is computed
    RK_(wait_exchange(h));
    trace_mark(h, 0, h->sc);
    if (h->cfg.overlap && lz > 2) {
        RK_((launch_step_range<SK, STRICT, GENERIC>(h, p, 1, 2, h->sc, lz - 1)));   // planes 1 and lz in one launch
        CK(cudaEventRecord(h->evB, h->sc));
        trace_mark(h, 1, h->sc);
        RK_((launch_step_range<SK, STRICT, GENERIC>(h, p, 2, lz - 2, h->sc)));
    } else {
        RK_((launch_step_range<SK, STRICT, GENERIC>(h, p, 1, lz, h->sc)));
        CK(cudaEventRecord(h->evB, h->sc));
        trace_mark(h, 1, h->sc);
    }
    trace_mark(h, 2, h->sc);
    CK(cudaStreamWaitEvent(h->sx, h->evB, 0));
    RK_(exchange_after_step(h, SK, written, h->sx));
    CK(cudaEventRecord(h->evX, h->sx));
    trace_mark(h, 3, h->sx);
    trace_next(h);
    h->exchange_pending = true;
    return 0;
}

template <bool STRICT, bool GENERIC>
static int step_dispatch(d3q19_handle *h, StepParams &p) {
    if (h->cfg.scheme == D3Q19_SCHEME_AB) {
        p.A = h->A; p.B = h->B;
        RK_((step_impl<STEP_AB, STRICT, GENERIC>(h, p, h->B)));
        double *t = h->A; h->A = h->B; h->B = t;
    } else if (h->phase == 0) {
        p.A = h->A; p.B = nullptr;
        RK_((step_impl<STEP_AA_EVEN, STRICT, GENERIC>(h, p, h->A)));
        h->phase = 1;
    } else {
        p.A = h->A; p.B = nullptr;
        RK_((step_impl<STEP_AA_ODD, STRICT, GENERIC>(h, p, h->A)));
        h->phase = 0;
    }
    h->n_steps++;
    return 0;
}

static int collide_stream_impl(d3q19_handle *h, int macro_mode, unsigned long long *rhoerr_bits) {
    if (macro_mode < 0 || macro_mode > 2) return fail("d3q19_collide_stream: bad macro_mode %d", macro_mode);
    RK_(ensure_mask(h));
    StepParams p;
    memset(&p, 0, sizeof p);
    p.g = h->g;
    p.mrt = Mrt{h->cfg.s1, h->cfg.s2, h->cfg.s4, h->cfg.s9, h->cfg.s10, h->cfg.s13, h->cfg.s16,
                h->cfg.omegepsl, h->cfg.omegepslj, h->cfg.omegxx};
    p.Fx = h->Fx; p.Fy = h->Fy; p.Fz = h->Fz;
    p.rho_shift = h->rho_shift;
    p.pf_ahead = h->pf_ahead;
    p.macro_mode = macro_mode;
    const bool generic = macro_mode != D3Q19_MACRO_MAIN || h->ffx || h->solid || h->rho_shift != 0.0;
    if (macro_mode != D3Q19_MACRO_MAIN) RK_(ensure_macro_arrays(h));
    p.rho = h->rho; p.ux = h->ux; p.uy = h->uy; p.uz = h->uz;
    p.ffx = h->ffx; p.ffy = h->ffy; p.ffz = h->ffz;
    p.solid = h->solid;
    p.rhoerr_bits = rhoerr_bits;
    const bool strict = h->cfg.math == D3Q19_MATH_STRICT;
    int rc;
    if (generic) rc = strict ? step_dispatch<true, true>(h, p) : step_dispatch<false, true>(h, p);
    else rc = strict ? step_dispatch<true, false>(h, p) : step_dispatch<false, false>(h, p);
    if (rc) return rc;
    if (macro_mode == D3Q19_MACRO_MAIN) h->rho_shift = 0.0;   // the shift lives for one collision (macrovar recomputes rho)
    h->shim.f_host_valid = false;       // also for the raw entry points (d3q19_run, d3q19_collide_stream, d3q19_prerelax)
    h->shim.macro_dev_valid = false;
    return 0;
}

extern "C" int d3q19_collide_stream(d3q19_handle *h, int32_t macro_mode) {
    CK(cudaSetDevice(h->cfg.device));
    return collide_stream_impl(h, macro_mode, nullptr);
}

extern "C" int d3q19_run(d3q19_handle *h, int32_t nsteps) {
    CK(cudaSetDevice(h->cfg.device));
    for (int i = 0; i < nsteps; ++i) RK_(collide_stream_impl(h, D3Q19_MACRO_MAIN, nullptr));
    return 0;
}

// ---- device-side initvel + initpop (initial.f90:75-147, :19-46) ------------------------------------------
extern "C" int d3q19_init_channel(d3q19_handle *h, double ustar, double ystar, double A9, double noise_amp,
                                  uint64_t seed, int32_t ivel) {
    CK(cudaSetDevice(h->cfg.device));
    RK_(wait_exchange(h));
    InitParams p;
    memset(&p, 0, sizeof p);
    p.g = h->g; p.A = h->A;
    p.nx = h->cfg.nx; p.ny = h->cfg.ny; p.nz = h->cfg.nz; p.globalz = h->cfg.globalz;
    p.ustar = ustar; p.ystar = ystar; p.A9 = A9; p.noise_amp = noise_amp;
    p.pi2 = 2.0 * (4.0 * atan(1.0));
    p.seed = seed; p.ivel = ivel;
    p.unstream = h->cfg.scheme == D3Q19_SCHEME_AB;
    k_init_channel<<<grid_nodes(h, h->g.lz), BLOCK_X, 0, h->sc>>>(p);
    CK(cudaGetLastError());
    h->n_other_kernels++;
    h->phase = 0;
    if (h->cfg.scheme == D3Q19_SCHEME_AB && h->cfg.nranks > 1) {
        // ghost planes of the post-collision storage, as after a step
        CK(cudaEventRecord(h->evB, h->sc));
        CK(cudaStreamWaitEvent(h->sx, h->evB, 0));
        RK_(exchange_after_step(h, STEP_AB, h->A, h->sx));
        CK(cudaEventRecord(h->evX, h->sx));
        h->exchange_pending = true;
    }
    h->shim.f_dev_valid = true;
    h->shim.f_host_valid = false;
    h->shim.macro_dev_valid = false;
    CK(cudaStreamSynchronize(h->sc));
    CK(cudaStreamSynchronize(h->sx));
    return halo_barrier(h);
}

// ---- macrovar / rhoupdat ---------------------------------------------------------------------------------
static int macro_launch(d3q19_handle *h, int rho_only, unsigned long long *rhoerr_bits = nullptr) {
    RK_(ensure_macro_arrays(h));
    RK_(ensure_mask(h));
    RK_(wait_exchange(h));
    MacroParams p;
    memset(&p, 0, sizeof p);
    p.g = h->g; p.A = h->A;
    p.rho = h->rho; p.ux = h->ux; p.uy = h->uy; p.uz = h->uz;
    p.Fx = h->Fx; p.Fy = h->Fy; p.Fz = h->Fz;
    p.ffx = h->ffx; p.ffy = h->ffy; p.ffz = h->ffz;
    p.solid = h->solid; p.isnodes = h->isn;
    p.ypglb = h->ypglb; p.wp = h->wp; p.omgp = h->omgp;
    p.rhopart = h->cfg.rhopart;
    p.ipart = h->cfg.ipart && h->isn && h->ypglb;
    p.ny = h->cfg.ny; p.nz = h->cfg.nz; p.globalz = h->cfg.globalz;
    p.rho_only = rho_only;
    p.rhoerr_bits = rhoerr_bits;
    const dim3 gr = grid_nodes(h, h->g.lz);
    switch (read_kind(h)) {
    case READ_DIRECT: k_macro<READ_DIRECT><<<gr, BLOCK_X, 0, h->sc>>>(p); break;
    case READ_PULL_NAT: k_macro<READ_PULL_NAT><<<gr, BLOCK_X, 0, h->sc>>>(p); break;
    default: k_macro<READ_PULL_SWAP><<<gr, BLOCK_X, 0, h->sc>>>(p); break;
    }
    CK(cudaGetLastError());
    h->n_other_kernels++;
    return 0;
}

extern "C" int d3q19_macrovar(d3q19_handle *h) {
    CK(cudaSetDevice(h->cfg.device));
    h->rho_shift = 0.0;     // macrovar recomputes rho from f: a pending avedensity shift is gone (collision.f90:418)
    RK_(macro_launch(h, 0));
    h->shim.macro_dev_valid = true;
    return 0;
}

extern "C" int d3q19_rhoupdat(d3q19_handle *h) {
    CK(cudaSetDevice(h->cfg.device));
    return macro_launch(h, 1);
}

extern "C" int d3q19_probe(d3q19_handle *h, int32_t ix, int32_t iy, int32_t iz, double out4[4]) {
    CK(cudaSetDevice(h->cfg.device));
    const Geom &g = h->g;
    if (ix < 1 || ix > g.lx || iy < 1 || iy > g.ly || iz < 1 || iz > g.lz) return fail("d3q19_probe: node out of range");
    RK_(wait_exchange(h));
    switch (read_kind(h)) {
    case READ_DIRECT: k_probe<READ_DIRECT><<<1, 1, 0, h->sc>>>(g, h->A, ix - 1, iy - 1, iz, h->Fx, h->Fy, h->Fz, h->ffx, h->ffy, h->ffz, h->scal); break;
    case READ_PULL_NAT: k_probe<READ_PULL_NAT><<<1, 1, 0, h->sc>>>(g, h->A, ix - 1, iy - 1, iz, h->Fx, h->Fy, h->Fz, h->ffx, h->ffy, h->ffz, h->scal); break;
    default: k_probe<READ_PULL_SWAP><<<1, 1, 0, h->sc>>>(g, h->A, ix - 1, iy - 1, iz, h->Fx, h->Fy, h->Fz, h->ffx, h->ffy, h->ffz, h->scal); break;
    }
    CK(cudaGetLastError());
    h->n_other_kernels++;
    CK(cudaMemcpyAsync(out4, h->scal, 4 * sizeof(double), cudaMemcpyDeviceToHost, h->sc));
    CK(cudaStreamSynchronize(h->sc));
    return 0;
}

// ---- avedensity (collision.f90:487-513) ----------------------------------------------------------------------
extern "C" int d3q19_avedensity(d3q19_handle *h, double *rhomean, int64_t *nfluid_total) {
    CK(cudaSetDevice(h->cfg.device));
    RK_(ensure_macro_arrays(h));
    RK_(ensure_mask(h));
    const Geom &g = h->g;
    const long long nrows = (long long)g.ly * g.lz;
    const int nblk = nrows < 1024 ? (int)nrows : 1024;             // a block per (y,z) row at most
    if (!h->red_d) {
        CK(cudaMalloc(&h->red_d, 1024 * sizeof(double)));
        CK(cudaMalloc(&h->red_c, (1024 + 8) * sizeof(long long)));
    }
    double *sum_dev = h->scal + 8;
    long long *cnt_dev = h->red_c + 1024;
    k_rho_partial<<<nblk, 256, 0, h->sc>>>(g.lx, g.xp, nrows, h->rho, h->solid, h->red_d, h->red_c);
    k_rho_final<<<1, 32, 0, h->sc>>>(nblk, h->red_d, h->red_c, sum_dev, cnt_dev);
    CK(cudaGetLastError());
    h->n_other_kernels += 2;
    if (h->cfg.nranks > 1) {           // MPI_ALLREDUCE x2, collision.f90:500-501
        NcclApi &n = nccl_api();
        NK(n.GroupStart());
        NK(n.AllReduce(sum_dev, sum_dev, 1, NCCL_FLOAT64, NCCL_SUM, h->comm, h->sc));
        NK(n.AllReduce(cnt_dev, cnt_dev, 1, NCCL_INT64, NCCL_SUM, h->comm, h->sc));
        NK(n.GroupEnd());
        h->n_nccl += 2;
    }
    double s = 0;
    long long c = 0;
    CK(cudaMemcpyAsync(&s, sum_dev, sizeof s, cudaMemcpyDeviceToHost, h->sc));
    CK(cudaMemcpyAsync(&c, cnt_dev, sizeof c, cudaMemcpyDeviceToHost, h->sc));
    CK(cudaStreamSynchronize(h->sc));
    if (c <= 0) return fail("d3q19_avedensity: no fluid nodes");
    const double mean = s / (double)c;                                 // collision.f90:503
    double *mean_dev = h->scal + 9;
    CK(cudaMemcpyAsync(mean_dev, &mean, sizeof mean, cudaMemcpyHostToDevice, h->sc));
    k_rho_shift<<<grid_nodes(h, g.lz), BLOCK_X, 0, h->sc>>>(g.lx, g.xp, h->rho, mean_dev);
    CK(cudaGetLastError());
    h->n_other_kernels++;
    CK(cudaStreamSynchronize(h->sc));
    h->rho_shift += mean;      // what the next collision_MRT sees in its rho array
    if (rhomean) *rhomean = mean;
    if (nfluid_total) *nfluid_total = c;
    return 0;
}

// ---- device pre-relaxation (main.f90:70-90) -----------------------------------------------------------------------
extern "C" int d3q19_prerelax(d3q19_handle *h, double tol, int32_t maxiter, int32_t *iters, double *rhoerrmax) {
    CK(cudaSetDevice(h->cfg.device));
    RK_(ensure_macro_arrays(h));
    unsigned long long *bits = reinterpret_cast<unsigned long long *>(h->scal + 16);
    int istep = 0;
    double err = 0.0;
    for (;;) {
        CK(cudaMemsetAsync(bits, 0, sizeof(unsigned long long), h->sc));
        RK_(collide_stream_impl(h, D3Q19_MACRO_PRERELAX, bits));      // rhoupdat fused into the collision
        if (h->cfg.nranks > 1) {                                       // MPI_ALLREDUCE MAX, main.f90:80
            NK(nccl_api().AllReduce(bits, bits, 1, NCCL_UINT64, NCCL_MAX, h->comm, h->sc));
            h->n_nccl++;
        }
        unsigned long long b = 0;
        CK(cudaMemcpyAsync(&b, bits, sizeof b, cudaMemcpyDeviceToHost, h->sc));
        CK(cudaStreamSynchronize(h->sc));
        memcpy(&err, &b, sizeof err);
        if (err <= tol || istep > maxiter) break;                     // main.f90:85
        istep++;
    }
    if (iters) *iters = istep;
    if (rhoerrmax) *rhoerrmax = err;
    return 0;
}

// ---- particles: masks and tables -------------------------------------------------------------------------------------
extern "C" int d3q19_set_solid_mask(d3q19_handle *h, const int32_t *ibnodes_ghosted, const int32_t *isnodes) {
    CK(cudaSetDevice(h->cfg.device));
    const Geom &g = h->g;
    if (!ibnodes_ghosted) {
        if (h->solid) { CK(cudaStreamSynchronize(h->sc)); cudaFree(h->solid); h->solid = nullptr; }
        return 0;
    }
    const size_t nb = h->nfield * sizeof(int32_t);
    if (!h->solid) CK(cudaMalloc(&h->solid, nb));
    const size_t nghost = (size_t)(g.lx + 2) * (g.ly + 2) * (g.lz + 2);
    if (isnodes && !h->isn) CK(cudaMalloc(&h->isn, nb));
    int32_t *tmp = nullptr;
    CK(cudaMalloc(&tmp, nghost * sizeof(int32_t)));
    cudaError_t e = cudaMemcpyAsync(tmp, ibnodes_ghosted, nghost * sizeof(int32_t), cudaMemcpyHostToDevice, h->sc);
    if (e == cudaSuccess) {
        k_field_unpack_i32<<<grid_nodes(h, g.lz), BLOCK_X, 0, h->sc>>>(g.lx, g.xp, h->solid, tmp, 1, g.ly);
        e = cudaGetLastError();
        h->n_other_kernels++;
    }
    if (e == cudaSuccess && isnodes) {
        e = cudaMemcpyAsync(tmp, isnodes, (size_t)g.lx * g.ly * g.lz * sizeof(int32_t), cudaMemcpyHostToDevice, h->sc);
        if (e == cudaSuccess) {
            k_field_unpack_i32<<<grid_nodes(h, g.lz), BLOCK_X, 0, h->sc>>>(g.lx, g.xp, h->isn, tmp, 0, g.ly);
            e = cudaGetLastError();
            h->n_other_kernels++;
        }
    }
    if (e == cudaSuccess) e = cudaStreamSynchronize(h->sc);
    cudaFree(tmp);                      // also on the error paths
    if (e != cudaSuccess) return fail("d3q19_set_solid_mask: %s", cudaGetErrorString(e));
    return 0;
}

// ---- particles: device-side bookkeeping (particles.cuh) --------------------------------------------------------------
static PartGeom part_geom(const d3q19_handle *h) {
    PartGeom pg;
    pg.g = h->g; pg.nx = h->cfg.nx; pg.ny = h->cfg.ny; pg.nz = h->cfg.nz; pg.globalz = h->cfg.globalz; pg.rad = h->pp.rad;
    return pg;
}

extern "C" int d3q19_particles_init(d3q19_handle *h, int32_t npart, const d3q19_particle_params *prm) {
    CK(cudaSetDevice(h->cfg.device));
    if (npart <= 0 || !prm || !(prm->rad > 0.0)) return fail("d3q19_particles_init: bad arguments");
    if (h->part_on) return fail("d3q19_particles_init: already initialised");
    if (!h->cfg.ipart) return fail("d3q19_particles_init: create the handle with ipart = 1");
    if (h->halo_on && h->halo_mode != D3Q19_HALO_PUT)
        return fail("d3q19_particles_init: with particles the faces travel by NCCL or by the copy engines (D3Q19_HALO_PUT), not by stores "
                    "inside the step kernel");
    if (!h->idx32) return fail("d3q19_particles_init: slab too large for 32-bit link indices");
    // a particle's bounding box (< 2 rad + 6 nodes wide) must be smaller than the periodic box: no node is visited twice
    if (2.0 * prm->rad + 7.0 > h->cfg.ny || 2.0 * prm->rad + 7.0 > h->cfg.nz) return fail("d3q19_particles_init: particle larger than the periodic box (2 rad + 7 <= ny, nz)");
    if (prm->rad > 600.0) return fail("d3q19_particles_init: rad %g: a particle's bounding box must hold fewer than 2^31 nodes", prm->rad);
    if (npart > 65535) return fail("d3q19_particles_init: at most 65535 particles (one grid row per particle)");
    h->pp = *prm;
    if (h->ypglb) { cudaFree(h->ypglb); cudaFree(h->wp); cudaFree(h->omgp); }
    if (h->solid) { cudaFree(h->solid); h->solid = nullptr; }
    if (h->isn) { cudaFree(h->isn); h->isn = nullptr; }
    h->npart = npart;
    const size_t tb = (size_t)3 * npart;
    CK(cudaMalloc(&h->pbuf, 10 * tb * sizeof(double)));
    CK(cudaMemsetAsync(h->pbuf, 0, 10 * tb * sizeof(double), h->sc));
    double *q = h->pbuf;
    h->ypglb = q; q += tb; h->ypmask = q; q += tb; h->wp = q; q += tb; h->omgp = q; q += tb; h->fHIp = q; q += tb;
    h->torqp = q; q += tb; h->flubp = q; q += tb; h->forcepp = q; q += tb; h->torqpp = q; q += tb; h->thetap = q;
    const size_t nown = (size_t)h->g.plane * (h->g.lz + 2);
    CK(cudaMalloc(&h->own, nown * sizeof(int32_t)));
    CK(cudaMemsetAsync(h->own, 0xFF, nown * sizeof(int32_t), h->sc));
    const double pi = 4.0 * atan(1.0);
    // every particle owns a segment of the link list: maxlink (all particles) / npart entries, by default 8 links per
    // surface node of a sphere of radius rad + 1 (cf. para.f90:371)
    h->links.cap = prm->maxlink > 0 ? (prm->maxlink + npart - 1) / npart : (long long)(8.0 * 4.0 * pi * (prm->rad + 1.0) * (prm->rad + 1.0)) + 64;
    h->maxlink = h->links.cap * npart;
    CK(cudaMalloc(&h->links.node, h->maxlink * sizeof(uint32_t)));
    CK(cudaMalloc(&h->links.dir, h->maxlink * sizeof(int32_t)));
    CK(cudaMalloc(&h->links.count, (size_t)npart * sizeof(unsigned long long)));
    CK(cudaMemsetAsync(h->links.count, 0, (size_t)npart * sizeof(unsigned long long), h->sc));
    CK(cudaMalloc(&h->pcnt, 4 * sizeof(unsigned long long)));
    CK(cudaMemsetAsync(h->pcnt, 0, 4 * sizeof(unsigned long long), h->sc));
    // a mask update uncovers at most a surface layer of every particle: the link capacity bounds it
    h->fill.cap = h->maxlink;
    CK(cudaMalloc(&h->fill.node, h->fill.cap * sizeof(uint32_t)));
    CK(cudaMalloc(&h->fill.part, h->fill.cap * sizeof(int32_t)));
    h->fill.count = h->pcnt + 1;
    h->part_rows = part_max_rows(prm->rad);
    const double volp = 4.0 / 3.0 * pi * prm->rad * prm->rad * prm->rad;      // para.f90:340
    h->amp = h->cfg.rhopart * volp;                                           // :341
    h->aip = 0.4 * h->amp * prm->rad * prm->rad;                              // :342
    h->solid = h->own + h->g.plane;
    h->isn = h->own + h->g.plane;
    h->part_on = true;
    h->links_valid = false; h->mask_built = false;
    CK(cudaStreamSynchronize(h->sc));
    return 0;
}

extern "C" int d3q19_set_particles(d3q19_handle *h, int32_t npart, const double *ypglb, const double *wp, const double *omgp) {
    CK(cudaSetDevice(h->cfg.device));
    if (npart <= 0 || !ypglb || !wp || !omgp) return fail("d3q19_set_particles: bad arguments");
    const size_t nb = (size_t)3 * npart * sizeof(double);
    if (h->part_on) {
        if (npart != h->npart) return fail("d3q19_set_particles: %d particles, initialised for %d", npart, h->npart);
        h->links_valid = false;
    } else if (npart != h->npart) {
        CK(cudaStreamSynchronize(h->sc));
        if (h->ypglb) { cudaFree(h->ypglb); cudaFree(h->wp); cudaFree(h->omgp); }
        CK(cudaMalloc(&h->ypglb, nb)); CK(cudaMalloc(&h->wp, nb)); CK(cudaMalloc(&h->omgp, nb));
        h->npart = npart;
    }
    CK(cudaMemcpyAsync(h->ypglb, ypglb, nb, cudaMemcpyHostToDevice, h->sc));
    CK(cudaMemcpyAsync(h->wp, wp, nb, cudaMemcpyHostToDevice, h->sc));
    CK(cudaMemcpyAsync(h->omgp, omgp, nb, cudaMemcpyHostToDevice, h->sc));
    if (h->part_on) {
        // the particles are PLACED, not moved: the mask starts afresh (nothing counts as uncovered, no refill)
        const size_t nown = (size_t)h->g.plane * (h->g.lz + 2);
        CK(cudaMemsetAsync(h->own, 0xFF, nown * sizeof(int32_t), h->sc));
        CK(cudaMemsetAsync(h->pcnt, 0, 4 * sizeof(unsigned long long), h->sc));
        h->mask_built = false;
    }
    CK(cudaStreamSynchronize(h->sc));
    return 0;
}

// the link counts live on the device (the step sequence never waits for them); the host asks when it must.
// k_beads_links drops the entries beyond a segment's capacity and k_beads_ibb skips them, which would let mass and momentum
// leak through the particle surface unnoticed: whoever fetches the counts checks them
static int fetch_link_counts(d3q19_handle *h, std::vector<unsigned long long> &cnt, const char *who) {
    cnt.assign((size_t)h->npart + 1, 0ull);
    CK(cudaMemcpyAsync(cnt.data(), h->links.count, (size_t)h->npart * sizeof(unsigned long long), cudaMemcpyDeviceToHost, h->sc));
    CK(cudaMemcpyAsync(&cnt[h->npart], h->pcnt + 1, sizeof(unsigned long long), cudaMemcpyDeviceToHost, h->sc));
    CK(cudaStreamSynchronize(h->sc));
    for (int p = 0; p < h->npart; ++p)
        if ((long long)cnt[p] > h->links.cap)
            return fail("%s: particle %d has %llu boundary links on this slab, its segment holds %lld -- links were dropped, the "
                        "populations are void (d3q19_particle_params.maxlink = links of ALL particles)", who, p + 1, cnt[p], h->links.cap);
    if ((long long)cnt[h->npart] > h->fill.cap)
        return fail("%s: the last move uncovered %llu nodes, the refill list holds %lld", who, cnt[h->npart], h->fill.cap);
    return 0;
}
static int fetch_nlink(d3q19_handle *h) {
    if (h->nlink >= 0) return 0;
    std::vector<unsigned long long> cnt;
    RK_(fetch_link_counts(h, cnt, "d3q19_beads_links"));
    long long n = 0;
    for (int p = 0; p < h->npart; ++p) n += (long long)cnt[p];
    h->nlink = n;
    return 0;
}
static int check_link_overflow(d3q19_handle *h, const char *who) {
    if (!h->part_on) return 0;
    std::vector<unsigned long long> cnt;
    return fetch_link_counts(h, cnt, who);
}

static inline dim3 sweep_grid(const d3q19_handle *h) {
    return dim3((unsigned)h->npart, (unsigned)((h->part_rows + PART_WARPS - 1) / PART_WARPS));
}

// beads_links: the solid mask follows the particle table, then the boundary-link list
extern "C" int d3q19_beads_links(d3q19_handle *h, int64_t *nlink_local) {
    CK(cudaSetDevice(h->cfg.device));
    if (!h->part_on) return fail("d3q19_beads_links: call d3q19_particles_init first");
    if (h->links_valid && h->mask_built) {
        // the particle table has not changed since the last build (d3q19_set_particles and d3q19_beads_move invalidate):
        // a second update would also forget which nodes the last move uncovered, which d3q19_beads_filling still needs
        if (nlink_local) { RK_(fetch_nlink(h)); *nlink_local = h->nlink; }
        return 0;
    }
    const PartGeom pg = part_geom(h);
    const dim3 gs = sweep_grid(h);
    const size_t tb = (size_t)3 * h->npart * sizeof(double);
    CK(cudaMemsetAsync(h->pcnt, 0, 2 * sizeof(unsigned long long), h->sc));          // refill list
    CK(cudaMemsetAsync(h->links.count, 0, (size_t)h->npart * sizeof(unsigned long long), h->sc));
    if (h->mask_built) {
        k_beads_uncover<<<gs, 32 * PART_WARPS, 0, h->sc>>>(pg, h->npart, h->ypmask, h->ypglb, h->own, h->fill);
        h->n_other_kernels++;
    }
    k_beads_cover<<<gs, 32 * PART_WARPS, 0, h->sc>>>(pg, h->npart, h->ypglb, h->own);
    CK(cudaMemcpyAsync(h->ypmask, h->ypglb, tb, cudaMemcpyDeviceToDevice, h->sc));
    h->mask_built = true;
    k_beads_links<<<gs, 32 * PART_WARPS, 0, h->sc>>>(pg, h->npart, h->ypglb, h->own, h->links);
    CK(cudaGetLastError());
    h->n_other_kernels += 2;
    h->links_valid = true;
    h->nlink = -1;                                                // known on the device; fetched on demand
    if (nlink_local) { RK_(fetch_nlink(h)); *nlink_local = h->nlink; }
    return 0;
}

// beads_collision: interpolated bounce-back on every link + hydrodynamic force and torque
extern "C" int d3q19_beads_collision(d3q19_handle *h) {
    CK(cudaSetDevice(h->cfg.device));
    if (!h->part_on || !h->links_valid) return fail("d3q19_beads_collision: no valid link list (d3q19_beads_links)");
    RK_(wait_exchange(h));
    const size_t tb = (size_t)3 * h->npart * sizeof(double);
    CK(cudaMemsetAsync(h->fHIp, 0, 2 * tb, h->sc));                // fHIp and torqp are adjacent
    {
        IbbParams P;
        P.pg = part_geom(h); P.S = h->A; P.own = h->own; P.L = h->links;
        P.ypglb = h->ypglb; P.wp = h->wp; P.omgp = h->omgp; P.rho0 = h->pp.rho0; P.fHIp = h->fHIp; P.torqp = h->torqp;
        // the counts are on the device: one grid row per particle, wide enough for a full segment; the blocks beyond a
        // segment's count leave at once
        const dim3 nb((unsigned)((h->links.cap + 127) / 128), (unsigned)h->npart);
        switch (read_kind(h)) {
        case READ_DIRECT: k_beads_ibb<READ_DIRECT><<<nb, 128, 0, h->sc>>>(P); break;
        case READ_PULL_NAT: k_beads_ibb<READ_PULL_NAT><<<nb, 128, 0, h->sc>>>(P); break;
        default: k_beads_ibb<READ_PULL_SWAP><<<nb, 128, 0, h->sc>>>(P); break;
        }
        CK(cudaGetLastError());
        h->n_other_kernels++;
    }
    if (h->cfg.nranks > 1) {                                       // force reduction over the slabs
        NK(nccl_api().AllReduce(h->fHIp, h->fHIp, (size_t)6 * h->npart, NCCL_FLOAT64, NCCL_SUM, h->comm, h->sc));
        h->n_nccl++;
    }
    h->shim.f_host_valid = false;
    return 0;
}

static int lubmove(d3q19_handle *h, int do_lub, int do_move) {
    LubParams lp = {h->pp.mingap, h->pp.mingap_w, h->pp.stf0, h->pp.stf1, h->pp.stf0_w, h->pp.stf1_w, h->pp.fscale};
    MoveParams M = {h->amp, h->aip, h->pp.gforce[0], h->pp.gforce[1], h->pp.gforce[2], h->fHIp, h->torqp, h->flubp,
                    h->forcepp, h->torqpp, h->ypglb, h->wp, h->omgp, h->thetap};
    const int nt = h->npart < 32 ? h->npart * 32 : 1024;          // a warp per particle for the partner loop
    if (do_lub && do_move && h->npart > 256) {
        // many particles: the O(npart^2) partner loop on several blocks, then the move as a launch of its own
        const int nb = (h->npart + 31) / 32;
        k_beads_lubmove<<<nb, 1024, 0, h->sc>>>(part_geom(h), h->npart, h->ypglb, lp, h->flubp, M, 1, 0);
        k_beads_lubmove<<<1, 1024, 0, h->sc>>>(part_geom(h), h->npart, h->ypglb, lp, h->flubp, M, 0, 1);
        h->n_other_kernels++;
    } else {
        k_beads_lubmove<<<1, nt, 0, h->sc>>>(part_geom(h), h->npart, h->ypglb, lp, h->flubp, M, do_lub, do_move);
    }
    CK(cudaGetLastError());
    h->n_other_kernels++;
    if (do_move) h->links_valid = false;
    return 0;
}

extern "C" int d3q19_beads_lubforce(d3q19_handle *h) {
    CK(cudaSetDevice(h->cfg.device));
    if (!h->part_on) return fail("d3q19_beads_lubforce: call d3q19_particles_init first");
    return lubmove(h, 1, 0);
}

extern "C" int d3q19_beads_move(d3q19_handle *h) {
    CK(cudaSetDevice(h->cfg.device));
    if (!h->part_on) return fail("d3q19_beads_move: call d3q19_particles_init first");
    return lubmove(h, 0, 1);
}

// Refill sources across a slab face: all 19 canonical populations of the neighbours' planes next to the faces (the ghost
// planes of the population array carry 5), so that the refill does not depend on the decomposition.  One launch gathers my
// two boundary planes, one NCCL group sends them to the neighbours (2 x 19 x plane fp64 per step and GPU).  `s`: the
// compute stream (stand-alone d3q19_beads_filling) or the second stream (d3q19_particle_step: the populations do not change
// between the bounce-back and the refill, so the exchange runs next to lubrication / move / mask / links).
static int fill_exchange(d3q19_handle *h, cudaStream_t s) {
    const Geom &g = h->g;
    const size_t cnt = (size_t)NPOP * g.plane;
    if (!h->fill_halo) CK(cudaMalloc(&h->fill_halo, 4 * cnt * sizeof(double)));
    double *send_up = h->fill_halo, *send_dn = send_up + cnt, *ghost_lo = send_dn + cnt, *ghost_hi = ghost_lo + cnt;
    const dim3 gp = grid_nodes(h, 2);             // blockIdx.z: 0 -> plane lz into send_up, 1 -> plane 1 into send_dn
    switch (read_kind(h)) {
    case READ_DIRECT: k_plane_gather<READ_DIRECT><<<gp, BLOCK_X, 0, s>>>(g, h->A, send_up, g.lz, send_dn, 1); break;
    case READ_PULL_NAT: k_plane_gather<READ_PULL_NAT><<<gp, BLOCK_X, 0, s>>>(g, h->A, send_up, g.lz, send_dn, 1); break;
    default: k_plane_gather<READ_PULL_SWAP><<<gp, BLOCK_X, 0, s>>>(g, h->A, send_up, g.lz, send_dn, 1); break;
    }
    CK(cudaGetLastError());
    const int up = (h->cfg.rank + 1) % h->cfg.nranks, dn = (h->cfg.rank + h->cfg.nranks - 1) % h->cfg.nranks;
    NcclApi &n = nccl_api();
    NK(n.GroupStart());
    NK(n.Send(send_up, cnt, NCCL_FLOAT64, up, h->comm, s));
    NK(n.Send(send_dn, cnt, NCCL_FLOAT64, dn, h->comm, s));
    NK(n.Recv(ghost_lo, cnt, NCCL_FLOAT64, dn, h->comm, s));
    NK(n.Recv(ghost_hi, cnt, NCCL_FLOAT64, up, h->comm, s));
    NK(n.GroupEnd());
    h->n_other_kernels += 1;
    h->n_nccl += 4;
    return 0;
}

// beads_filling: populations of the nodes the last move uncovered (needs the rebuilt mask)
extern "C" int d3q19_beads_filling(d3q19_handle *h, int64_t *nfilled) {
    CK(cudaSetDevice(h->cfg.device));
    if (!h->part_on || !h->links_valid) return fail("d3q19_beads_filling: rebuild the mask first (d3q19_beads_links)");
    RK_(wait_exchange(h));
    CK(cudaMemsetAsync(h->pcnt + 2, 0, sizeof(unsigned long long), h->sc));
    FillParams P;
    P.pg = part_geom(h); P.S = h->A; P.own = h->own; P.F = h->fill;
    P.ypglb = h->ypglb; P.wp = h->wp; P.omgp = h->omgp; P.nfilled = h->pcnt + 2;
    P.ghost_lo = P.ghost_hi = nullptr;
    if (h->cfg.nranks > 1) {
        // source nodes across a slab face: the neighbours' boundary planes (fill_exchange); d3q19_particle_step started the
        // exchange on the second stream right after the bounce-back, next to the bookkeeping kernels
        if (h->fill_xchg_pending) { CK(cudaStreamWaitEvent(h->sc, h->evF, 0)); h->fill_xchg_pending = false; }
        else RK_(fill_exchange(h, h->sc));
        const size_t cnt = (size_t)NPOP * h->g.plane;
        P.ghost_lo = h->fill_halo + 2 * cnt; P.ghost_hi = h->fill_halo + 3 * cnt;
    }
    // the list length is on the device: a fixed grid strides over it (one move uncovers a thin layer: a few dozen nodes per particle)
    const unsigned nb = (unsigned)(h->fill.cap < 148 * 8 * 128 ? (h->fill.cap + 127) / 128 : 148 * 8);
    switch (read_kind(h)) {
    case READ_DIRECT: k_beads_fill<READ_DIRECT><<<nb, 128, 0, h->sc>>>(P); break;
    case READ_PULL_NAT: k_beads_fill<READ_PULL_NAT><<<nb, 128, 0, h->sc>>>(P); break;
    default: k_beads_fill<READ_PULL_SWAP><<<nb, 128, 0, h->sc>>>(P); break;
    }
    // the list is consumed: a second call before the next move fills nothing
    CK(cudaMemsetAsync(h->pcnt + 1, 0, sizeof(unsigned long long), h->sc));
    CK(cudaGetLastError());
    h->n_other_kernels++;
    if (nfilled) {
        unsigned long long n = 0;
        CK(cudaMemcpyAsync(&n, h->pcnt + 2, sizeof n, cudaMemcpyDeviceToHost, h->sc));
        CK(cudaStreamSynchronize(h->sc));
        *nfilled = (int64_t)n;
    }
    h->shim.f_host_valid = false;
    return 0;
}

// one particle-laden time step in the order of the reference's timers (var_inc.f90:166-168)
extern "C" int d3q19_particle_step(d3q19_handle *h, int32_t move) {
    CK(cudaSetDevice(h->cfg.device));
    if (!h->part_on) return fail("d3q19_particle_step: call d3q19_particles_init first");
    if (!h->links_valid) RK_(d3q19_beads_links(h, nullptr));
    RK_(collide_stream_impl(h, D3Q19_MACRO_MAIN, nullptr));       // fluid nodes only (solid nodes skipped)
    RK_(d3q19_beads_collision(h));
    if (move) {
        if (h->cfg.nranks > 1) {
            // the refill's sources travel while the particles move and the mask and the links are rebuilt
            if (!h->evP) { CK(cudaEventCreateWithFlags(&h->evP, cudaEventDisableTiming)); CK(cudaEventCreateWithFlags(&h->evF, cudaEventDisableTiming)); }
            CK(cudaEventRecord(h->evP, h->sc));                     // after the bounce-back (which waited for the faces)
            CK(cudaStreamWaitEvent(h->sx, h->evP, 0));
            RK_(fill_exchange(h, h->sx));
            CK(cudaEventRecord(h->evF, h->sx));
            h->fill_xchg_pending = true;
        }
        RK_(lubmove(h, 1, 1));                                      // beads_lubforce + beads_move
        RK_(d3q19_beads_links(h, nullptr));
        RK_(d3q19_beads_filling(h, nullptr));
    }
    return 0;
}

extern "C" int d3q19_get_particles(d3q19_handle *h, double *ypglb, double *wp, double *omgp, double *fHIp, double *torqp) {
    CK(cudaSetDevice(h->cfg.device));
    if (!h->part_on) return fail("d3q19_get_particles: call d3q19_particles_init first");
    const size_t tb = (size_t)3 * h->npart * sizeof(double);
    const double *src[5] = {h->ypglb, h->wp, h->omgp, h->fHIp, h->torqp};
    double *dst[5] = {ypglb, wp, omgp, fHIp, torqp};
    for (int i = 0; i < 5; ++i) if (dst[i]) CK(cudaMemcpyAsync(dst[i], src[i], tb, cudaMemcpyDeviceToHost, h->sc));
    CK(cudaStreamSynchronize(h->sc));
    return check_link_overflow(h, "d3q19_get_particles");
}

// the link list of this slab, the particles' segments concatenated: global 1-based node coordinates, direction, particle, q
// (inside a particle's segment the order depends on the run: sort before comparing)
extern "C" int d3q19_get_links(d3q19_handle *h, int64_t capacity, int32_t *x, int32_t *y, int32_t *z, int32_t *ip,
                               int32_t *part, double *q, int64_t *nlink) {
    CK(cudaSetDevice(h->cfg.device));
    if (!h->part_on || !h->links_valid) return fail("d3q19_get_links: no valid link list");
    std::vector<unsigned long long> cnt;
    RK_(fetch_link_counts(h, cnt, "d3q19_get_links"));
    std::vector<long long> off((size_t)h->npart + 1, 0);
    for (int p = 0; p < h->npart; ++p) off[p + 1] = off[p] + (long long)cnt[p];
    h->nlink = off[h->npart];
    if (nlink) *nlink = h->nlink;
    if (capacity < h->nlink) return fail("d3q19_get_links: capacity %lld < %lld links", (long long)capacity, h->nlink);
    if (h->nlink == 0) return 0;
    const size_t n = (size_t)h->nlink;
    char *tmp = nullptr;
    const size_t bytes = (size_t)(h->npart + 1) * sizeof(long long) + 5 * n * sizeof(int32_t) + n * sizeof(double) + 64;
    CK(cudaMalloc(&tmp, bytes));
    double *dq = reinterpret_cast<double *>(tmp);                                   // 8-byte items first
    long long *doff = reinterpret_cast<long long *>(dq + n);
    int32_t *di = reinterpret_cast<int32_t *>(doff + h->npart + 1);
    cudaError_t e = cudaMemcpyAsync(doff, off.data(), (size_t)(h->npart + 1) * sizeof(long long), cudaMemcpyHostToDevice, h->sc);
    if (e == cudaSuccess) {
        const dim3 gr((unsigned)((h->links.cap + 127) / 128), (unsigned)h->npart);
        k_links_export<<<gr, 128, 0, h->sc>>>(part_geom(h), h->links, h->ypglb, doff, di, di + n, di + 2 * n, di + 3 * n, di + 4 * n, dq);
        e = cudaGetLastError();
    }
    void *dst[5] = {x, y, z, ip, part};
    for (int i = 0; i < 5 && e == cudaSuccess; ++i)
        e = cudaMemcpyAsync(dst[i], di + (size_t)i * n, n * sizeof(int32_t), cudaMemcpyDeviceToHost, h->sc);
    if (e == cudaSuccess) e = cudaMemcpyAsync(q, dq, n * sizeof(double), cudaMemcpyDeviceToHost, h->sc);
    if (e == cudaSuccess) e = cudaStreamSynchronize(h->sc);
    cudaFree(tmp);                      // also on the error paths
    if (e != cudaSuccess) return fail("d3q19_get_links: %s", cudaGetErrorString(e));
    return 0;
}

// owner mask of the local slab without ghosts, (lx,ly,lz): particle id (1-based) or -1 (= isnodes; ibnodes = sign)
extern "C" int d3q19_get_mask(d3q19_handle *h, int32_t *own_lx_ly_lz) {
    CK(cudaSetDevice(h->cfg.device));
    if (!h->part_on || !h->mask_built) return fail("d3q19_get_mask: no mask yet");
    const Geom &g = h->g;
    CK(cudaMemcpy2DAsync(own_lx_ly_lz, (size_t)g.lx * sizeof(int32_t), h->own + g.plane, (size_t)g.xp * sizeof(int32_t),
                         (size_t)g.lx * sizeof(int32_t), (size_t)g.ly * g.lz, cudaMemcpyDeviceToHost, h->sc));
    CK(cudaStreamSynchronize(h->sc));
    // "uncovered by q in the last update" (-(q+2), particles.cuh) is fluid: the caller sees isnodes = -1 there
    const size_t nn = (size_t)g.lx * g.ly * g.lz;
    for (size_t i = 0; i < nn; ++i) if (own_lx_ly_lz[i] < 0) own_lx_ly_lz[i] = -1;
    return 0;
}

// ---- profiles (statistc, saveload.f90:1241-1300) -------------------------------------------------------------------------
static int profiles_impl(d3q19_handle *h, double *out, int nrows_out);
extern "C" int d3q19_profiles(d3q19_handle *h, double *out) { return profiles_impl(h, out, 11); }
// statistc2 (saveload.f90:1348-1502): the same masked sums plus, as row 11, the fluid-node count of each x-plane
extern "C" int d3q19_profiles2(d3q19_handle *h, double *out) { return profiles_impl(h, out, 12); }
static int profiles_impl(d3q19_handle *h, double *out, int nrows_out) {
    CK(cudaSetDevice(h->cfg.device));
    const Geom &g = h->g;
    RK_(ensure_mask(h));
    RK_(wait_exchange(h));
    const long long nrows = (long long)g.ly * g.lz;
    if (!h->prof_partial) {
        int chunks = 296;                                     // 2 per SM
        if (chunks > nrows) chunks = (int)nrows;
        h->prof_rows = (int)((nrows + chunks - 1) / chunks);
        h->prof_chunks = (int)((nrows + h->prof_rows - 1) / h->prof_rows);
        CK(cudaMalloc(&h->prof_partial, (size_t)h->prof_chunks * NPROF * g.lx * sizeof(double)));
        CK(cudaMalloc(&h->prof_out, (size_t)NPROF * g.lx * sizeof(double)));
    }
    const dim3 gr((unsigned)((g.lx + BLOCK_X - 1) / BLOCK_X), (unsigned)h->prof_chunks);
    switch (read_kind(h)) {
    case READ_DIRECT: k_profiles<READ_DIRECT><<<gr, BLOCK_X, 0, h->sc>>>(g, h->A, h->Fx, h->Fy, h->Fz, h->ffx, h->ffy, h->ffz, h->solid, h->prof_rows, h->prof_partial); break;
    case READ_PULL_NAT: k_profiles<READ_PULL_NAT><<<gr, BLOCK_X, 0, h->sc>>>(g, h->A, h->Fx, h->Fy, h->Fz, h->ffx, h->ffy, h->ffz, h->solid, h->prof_rows, h->prof_partial); break;
    default: k_profiles<READ_PULL_SWAP><<<gr, BLOCK_X, 0, h->sc>>>(g, h->A, h->Fx, h->Fy, h->Fz, h->ffx, h->ffy, h->ffz, h->solid, h->prof_rows, h->prof_partial); break;
    }
    const dim3 g2((unsigned)((g.lx + BLOCK_X - 1) / BLOCK_X), NPROF);
    k_profiles_final<<<g2, BLOCK_X, 0, h->sc>>>(g.lx, h->prof_chunks, h->prof_partial, h->prof_out);
    CK(cudaGetLastError());
    h->n_other_kernels += 2;
    if (h->cfg.nranks > 1) {
        NK(nccl_api().AllReduce(h->prof_out, h->prof_out, (size_t)NPROF * g.lx, NCCL_FLOAT64, NCCL_SUM, h->comm, h->sc));
        h->n_nccl++;
    }
    CK(cudaMemcpyAsync(out, h->prof_out, (size_t)nrows_out * g.lx * sizeof(double), cudaMemcpyDeviceToHost, h->sc));
    CK(cudaStreamSynchronize(h->sc));
    return 0;
}

// ---- diag (saveload.f90:1507-1676) on the device ------------------------------------------------------------------------
// out[14]: vmax, imout, jmout, kmout (global 1-based), umean, vmean, wmean, urms, vrms, wrms (all / ustar),
//          volf, rhomax, rhomin, nfluid -- what the reference writes to diag.dat (format 260).
extern "C" int d3q19_diag(d3q19_handle *h, double ustar, double *out14) {
    CK(cudaSetDevice(h->cfg.device));
    const Geom &g = h->g;
    RK_(ensure_mask(h));
    RK_(wait_exchange(h));
    const long long nrows = (long long)g.ly * g.lz;
    int chunks = 296;
    if (chunks > nrows) chunks = (int)nrows;
    const int rows = (int)((nrows + chunks - 1) / chunks);
    chunks = (int)((nrows + rows - 1) / rows);
    const dim3 gr((unsigned)((g.lx + BLOCK_X - 1) / BLOCK_X), (unsigned)chunks);
    const int npartial = (int)(gr.x * gr.y);
    if (!h->diag_partial || h->diag_npartial < npartial) {
        if (h->diag_partial) { CK(cudaStreamSynchronize(h->sc)); cudaFree(h->diag_partial); h->diag_partial = nullptr; }
        CK(cudaMalloc(&h->diag_partial, ((size_t)npartial + 1) * NDIAG * sizeof(double)));
        h->diag_npartial = npartial;
    }
    double *partial = h->diag_partial;
    double *res = partial + (size_t)npartial * NDIAG;
    switch (read_kind(h)) {
    case READ_DIRECT: k_diag<READ_DIRECT><<<gr, BLOCK_X, 0, h->sc>>>(g, h->A, h->Fx, h->Fy, h->Fz, h->ffx, h->ffy, h->ffz, h->solid, rows, partial); break;
    case READ_PULL_NAT: k_diag<READ_PULL_NAT><<<gr, BLOCK_X, 0, h->sc>>>(g, h->A, h->Fx, h->Fy, h->Fz, h->ffx, h->ffy, h->ffz, h->solid, rows, partial); break;
    default: k_diag<READ_PULL_SWAP><<<gr, BLOCK_X, 0, h->sc>>>(g, h->A, h->Fx, h->Fy, h->Fz, h->ffx, h->ffy, h->ffz, h->solid, rows, partial); break;
    }
    k_diag_final<<<1, 32, 0, h->sc>>>(npartial, partial, res);
    CK(cudaGetLastError());
    h->n_other_kernels += 2;
    double loc[NDIAG];
    CK(cudaMemcpyAsync(loc, res, sizeof loc, cudaMemcpyDeviceToHost, h->sc));
    CK(cudaStreamSynchronize(h->sc));
    // this rank's maximum in global coordinates (saveload.f90:1566-1574)
    double mine[4] = {loc[7], 0, 0, 0};
    if (loc[8] >= 0.0) {
        const long long li = (long long)loc[8];
        mine[1] = (double)(li % g.lx + 1);
        mine[2] = (double)((li / g.lx) % g.ly + 1);
        mine[3] = (double)(li / ((long long)g.lx * g.ly) + 1 + h->cfg.globalz);
    }
    double sums[7] = {loc[0], loc[1], loc[2], loc[3], loc[4], loc[5], loc[6]};
    double rmax = loc[9], rmin = loc[10];
    std::vector<double> all((size_t)4 * h->cfg.nranks, 0.0);
    if (h->cfg.nranks > 1) {
        // MPI_ALLREDUCE of the sums, the per-rank maxima and rho extrema (saveload.f90:1544-1551,1582-1586,1617-1622)
        const int nr = h->cfg.nranks;
        if (!h->diag_red) CK(cudaMalloc(&h->diag_red, (size_t)(9 + 4 + 4 * nr) * sizeof(double)));
        double *d = h->diag_red;
        double pack[13] = {sums[0], sums[1], sums[2], sums[3], sums[4], sums[5], sums[6], rmax, -rmin, mine[0], mine[1], mine[2], mine[3]};
        CK(cudaMemcpyAsync(d, pack, sizeof pack, cudaMemcpyHostToDevice, h->sc));
        NcclApi &n = nccl_api();
        NK(n.GroupStart());
        NK(n.AllReduce(d, d, 7, NCCL_FLOAT64, NCCL_SUM, h->comm, h->sc));
        NK(n.AllReduce(d + 7, d + 7, 2, NCCL_FLOAT64, NCCL_MAX, h->comm, h->sc));
        NK(n.AllGather(d + 9, d + 13, 4, NCCL_FLOAT64, h->comm, h->sc));
        NK(n.GroupEnd());
        h->n_nccl += 3;
        CK(cudaMemcpyAsync(pack, d, 9 * sizeof(double), cudaMemcpyDeviceToHost, h->sc));
        CK(cudaMemcpyAsync(all.data(), d + 13, (size_t)4 * nr * sizeof(double), cudaMemcpyDeviceToHost, h->sc));
        CK(cudaStreamSynchronize(h->sc));
        for (int q = 0; q < 7; ++q) sums[q] = pack[q];
        rmax = pack[7]; rmin = -pack[8];
    } else {
        for (int q = 0; q < 4; ++q) all[q] = mine[q];
    }
    double vmax = 0.0, im = 0, jm = 0, km = 0;
    for (int r = 0; r < h->cfg.nranks; ++r)                       // first rank with a strictly larger value (saveload.f90:1649-1656)
        if (all[4 * r] > vmax) { vmax = all[4 * r]; im = all[4 * r + 1]; jm = all[4 * r + 2]; km = all[4 * r + 3]; }
    const double nf = sums[0];
    if (nf <= 0.0) return fail("d3q19_diag: no fluid nodes");
    const double um = sums[1] / nf, vm = sums[2] / nf, wm = sums[3] / nf;
    out14[0] = vmax; out14[1] = im; out14[2] = jm; out14[3] = km;
    out14[4] = um / ustar; out14[5] = vm / ustar; out14[6] = wm / ustar;
    out14[7] = sqrt(sums[4] / nf - um * um) / ustar;
    out14[8] = sqrt(sums[5] / nf - vm * vm) / ustar;
    out14[9] = sqrt(sums[6] / nf - wm * wm) / ustar;
    out14[10] = 1.0 - nf / ((double)h->cfg.nx * h->cfg.ny * h->cfg.nz);
    out14[11] = rmax; out14[12] = rmin; out14[13] = nf;
    return 0;
}

// ---- measurement ---------------------------------------------------------------------------------------------------
extern "C" int d3q19_timer_start(d3q19_handle *h) {
    CK(cudaSetDevice(h->cfg.device));
    RK_(wait_exchange(h));
    CK(cudaEventRecord(h->t0, h->sc));
    return 0;
}

extern "C" int d3q19_timer_stop(d3q19_handle *h, float *ms) {
    CK(cudaSetDevice(h->cfg.device));
    RK_(wait_exchange(h));     // the last exchange is part of the last step
    CK(cudaEventRecord(h->t1, h->sc));
    CK(cudaEventSynchronize(h->t1));
    CK(cudaEventElapsedTime(ms, h->t0, h->t1));
    return 0;
}

// ---- per-step timeline (development aid: where does a multi-GPU step spend its time) -----------------------------
extern "C" int d3q19_trace_enable(d3q19_handle *h, int32_t max_steps) {
    CK(cudaSetDevice(h->cfg.device));
    if (h->trace_ev) {
        for (int i = 0; i < 4 * h->trace_cap; ++i) cudaEventDestroy(h->trace_ev[i]);
        delete[] h->trace_ev;
        h->trace_ev = nullptr;
    }
    h->trace_cap = h->trace_n = 0;
    if (max_steps <= 0) return 0;
    h->trace_ev = new (std::nothrow) cudaEvent_t[4 * (size_t)max_steps];
    if (!h->trace_ev) return fail("d3q19_trace_enable: out of host memory");
    for (int i = 0; i < 4 * max_steps; ++i) CK(cudaEventCreate(&h->trace_ev[i]));
    h->trace_cap = max_steps;
    return 0;
}

// out[4*s + k] = milliseconds from the first mark of step 0 to mark k of step s (k: 0 before the boundary launch,
// 1 after it, 2 after the interior launch, 3 after the exchange); returns the number of recorded steps
extern "C" int d3q19_trace_fetch(d3q19_handle *h, int32_t *nsteps, float *out, int32_t capacity_steps) {
    CK(cudaSetDevice(h->cfg.device));
    RK_(d3q19_sync(h));
    const int n = h->trace_n < capacity_steps ? h->trace_n : capacity_steps;
    for (int s = 0; s < n; ++s)
        for (int k = 0; k < 4; ++k) CK(cudaEventElapsedTime(&out[4 * s + k], h->trace_ev[0], h->trace_ev[4 * s + k]));
    *nsteps = n;
    h->trace_n = 0;
    return 0;
}

extern "C" int d3q19_get_counters(d3q19_handle *h, int64_t out[8]) {
    out[0] = h->n_step_kernels; out[1] = h->n_other_kernels; out[2] = h->n_nccl; out[3] = h->n_steps;
    out[4] = (int64_t)((size_t)NPOP * h->g.slab * sizeof(double) * (h->cfg.scheme == D3Q19_SCHEME_AB ? 2 : 1));
    out[5] = h->phase; out[6] = h->g.xp; out[7] = h->cfg.scheme;
    return 0;
}

// ---- shim state machine (SURVEY.md section 8(b) "state coherence") ----------------------------------------------------
// The driver's arrays are ordinary pageable memory (Fortran ALLOCATE, para.f90:418-503).  Copies from and to
// pageable memory are staged by the CUDA driver at a fraction of the PCIe rate, and they are what the end-to-end
// time of the intact driver consists of beyond the step kernels (f up once, rho,u down on output steps), so
// large arrays are page-locked in place.  Arrays below 64 MB (tests) are left alone; memory that is pinned
// already (bench.py, torch) reports so and is left alone too.
static void shim_unpin(d3q19_handle *h) {
    for (int i = 0; i < h->shim.npinned; ++i)
        if (h->shim.pinned[i] && cudaHostUnregister(h->shim.pinned[i]) != cudaSuccess) cudaGetLastError();
    h->shim.npinned = 0;
}
static void shim_pin(d3q19_handle *h) {
    const size_t n = (size_t)h->g.lx * h->g.ly * h->g.lz * sizeof(double);
    if (n < ((size_t)64 << 20)) return;
    cudaSetDevice(h->cfg.device);
    void *ptr[5] = {h->shim.a.f, h->shim.a.rho, h->shim.a.ux, h->shim.a.uy, h->shim.a.uz};
    const size_t bytes[5] = {NPOP * n, n, n, n, n};
    for (int i = 0; i < 5; ++i) {
        if (!ptr[i]) continue;
        if (cudaHostRegister(ptr[i], bytes[i], cudaHostRegisterDefault) == cudaSuccess) h->shim.pinned[h->shim.npinned++] = ptr[i];
        else cudaGetLastError();          // already pinned by the caller, or not lockable: copies still work, only slower
    }
}

extern "C" int d3q19_shim_bind(d3q19_handle *h, const d3q19_shim_arrays *a) {
    if (!a || !a->f) return fail("d3q19_shim_bind: f is required");
    shim_unpin(h);
    h->shim.a = *a;
    shim_pin(h);
    h->shim.bound = true;
    h->shim.f_host_valid = true;
    h->shim.f_dev_valid = false;
    h->shim.macro_dev_valid = false;
    h->shim.u_dev_loaded = false;
    h->shim.next_mode = D3Q19_MACRO_MAIN;
    h->shim.collided_since_macrovar = false;
    h->shim.in_prerelax = false;
    h->shim.prerelax_iter = 0;
    h->shim.prerelax_err = 0.0;
    if (h->shim.a.prerelax_maxiter <= 0) h->shim.a.prerelax_maxiter = 15000;   // main.f90:85
    return 0;
}

// the same binding with the arrays as plain by-reference arguments: what a Fortran caller passes
// without C_LOC (var_inc's arrays have no TARGET attribute)
extern "C" int d3q19_shim_bind_arrays(d3q19_handle *h, double *f, double *rho, double *ux, double *uy, double *uz,
                                      double *force_realx, double *force_realy, double *force_realz, int32_t *ibnodes,
                                      int32_t *isnodes, int32_t has_isnodes, int32_t ndiag, int32_t nflowout,
                                      int32_t nsteps_total, int32_t istep0, int32_t ntime, int32_t prerelax_maxiter,
                                      double rhoepsl) {
    d3q19_shim_arrays a;
    memset(&a, 0, sizeof a);
    a.f = f; a.rho = rho; a.ux = ux; a.uy = uy; a.uz = uz;
    a.force_realx = force_realx; a.force_realy = force_realy; a.force_realz = force_realz;
    a.ibnodes = ibnodes; a.isnodes = has_isnodes ? isnodes : nullptr;
    a.ndiag = ndiag; a.nflowout = nflowout; a.nsteps_total = nsteps_total; a.istep0 = istep0;
    a.ntime = ntime; a.prerelax_maxiter = prerelax_maxiter; a.rhoepsl = rhoepsl;
    return d3q19_shim_bind(h, &a);
}

extern "C" int d3q19_shim_set_schedule(d3q19_handle *h, int32_t ndiag, int32_t nflowout, int32_t nsteps_total, int32_t istep0) {
    if (!h->shim.bound) return fail("d3q19_shim_*: call d3q19_shim_bind first");
    h->shim.a.ndiag = ndiag; h->shim.a.nflowout = nflowout;
    h->shim.a.nsteps_total = nsteps_total; h->shim.a.istep0 = istep0;
    return 0;
}

static int shim_f_on_device(d3q19_handle *h) {
    Shim &s = h->shim;
    if (!s.bound) return fail("d3q19_shim_*: call d3q19_shim_bind first");
    if (!s.f_dev_valid) {
        RK_(d3q19_upload_f(h, s.a.f));
        if (h->cfg.ipart && s.a.ibnodes) RK_(d3q19_set_solid_mask(h, s.a.ibnodes, s.a.isnodes));
        s.f_dev_valid = true;
    }
    return 0;
}

extern "C" int d3q19_shim_forcing(d3q19_handle *h, double force_in_y, double force_mag) {
    Shim &s = h->shim;
    const double fy = force_in_y * force_mag;                 // collision.f90:523
    if (s.bound && s.a.force_realx && s.a.force_realy && s.a.force_realz) {
        const size_t n = (size_t)h->g.lx * h->g.ly * h->g.lz;  // keep the host arrays what FORCING makes them
        for (size_t i = 0; i < n; ++i) { s.a.force_realx[i] = 0.0; s.a.force_realy[i] = fy; s.a.force_realz[i] = 0.0; }
    }
    return d3q19_set_force_uniform(h, 0.0, fy, 0.0);
}

extern "C" int d3q19_shim_rhoupdat(d3q19_handle *h) {
    Shim &s = h->shim;
    RK_(shim_f_on_device(h));
    if (!s.u_dev_loaded) {                                     // u stays frozen at the driver's values
        RK_(d3q19_set_macro(h, s.a.rho, s.a.ux, s.a.uy, s.a.uz));
        s.u_dev_loaded = true;
        s.prerelax_iter = 0;
    }
    // rho = sum f, and the number main.f90:79-80 is about to form from the host arrays
    unsigned long long *bits = reinterpret_cast<unsigned long long *>(h->scal + 17);
    CK(cudaMemsetAsync(bits, 0, sizeof(unsigned long long), h->sc));
    RK_(macro_launch(h, 1, bits));
    if (h->cfg.nranks > 1) {                                   // MPI_ALLREDUCE MAX, main.f90:80
        NK(nccl_api().AllReduce(bits, bits, 1, NCCL_UINT64, NCCL_MAX, h->comm, h->sc));
        h->n_nccl++;
    }
    unsigned long long b = 0;
    CK(cudaMemcpyAsync(&b, bits, sizeof b, cudaMemcpyDeviceToHost, h->sc));
    RK_(d3q19_download_macro(h, s.a.rho, nullptr, nullptr, nullptr));   // synchronises sc
    memcpy(&s.prerelax_err, &b, sizeof b);
    s.in_prerelax = true;
    s.next_mode = D3Q19_MACRO_EXTERNAL;                        // collision reads the arrays as they now are
    return 0;
}

extern "C" int d3q19_shim_prerelax_state(d3q19_handle *h, double *rhoerrmax, int32_t *iteration) {
    if (rhoerrmax) *rhoerrmax = h->shim.prerelax_err;
    if (iteration) *iteration = h->shim.prerelax_iter;
    return 0;
}

extern "C" int d3q19_shim_collision_mrt(d3q19_handle *h) {
    Shim &s = h->shim;
    RK_(shim_f_on_device(h));
    const bool main_loop = s.next_mode == D3Q19_MACRO_MAIN;
    RK_(d3q19_collide_stream(h, s.next_mode));
    s.next_mode = D3Q19_MACRO_MAIN;
    s.f_host_valid = false;
    s.macro_dev_valid = false;
    if (main_loop) s.collided_since_macrovar = true;
    if (s.in_prerelax) {
        // main.f90:85: the driver leaves the loop after THIS iteration iff the test below holds;
        // what follows is saveinitflow (main.f90:101), which writes the host f
        s.in_prerelax = false;
        const bool leaving = s.prerelax_err <= s.a.rhoepsl || s.prerelax_iter > s.a.prerelax_maxiter;
        s.prerelax_iter++;
        if (leaving) RK_(d3q19_shim_sync_f_to_host(h));
    }
    return 0;
}

extern "C" int d3q19_shim_macrovar(d3q19_handle *h, int32_t istep) {
    Shim &s = h->shim;
    RK_(shim_f_on_device(h));
    const d3q19_shim_arrays &a = s.a;
    // A macrovar that does not follow a main-loop collision_MRT is one of the initialisation
    // calls (main.f90:102,136): its output is always read.  Inside the loop the host arrays are
    // refreshed only on the steps where the intact driver reads them.
    const bool wanted = !s.collided_since_macrovar
                        || (a.ndiag > 0 && istep % a.ndiag == 0)            // diag, main.f90:171
                        || (a.nflowout > 0 && istep % a.nflowout == 0)      // outputflow, main.f90:184
                        || (a.ntime > 0 && istep % a.ntime == 0)            // the loop may exit here, main.f90:197-206
                        || istep >= a.istep0 + a.nsteps_total               // probe after the loop, main.f90:221
                        || (h->cfg.ipart && istep % 100 == 0);              // avedensity, main.f90:163
    s.collided_since_macrovar = false;
    s.u_dev_loaded = false;
    if (!wanted) return 0;                                                  // the next collision recomputes moments in registers
    RK_(d3q19_macrovar(h));
    s.macro_dev_valid = true;
    return d3q19_download_macro(h, a.rho, a.ux, a.uy, a.uz);
}

extern "C" int d3q19_shim_avedensity(d3q19_handle *h) {
    Shim &s = h->shim;
    RK_(shim_f_on_device(h));
    if (!s.macro_dev_valid) { RK_(d3q19_macrovar(h)); s.macro_dev_valid = true; }
    RK_(d3q19_avedensity(h, nullptr, nullptr));
    return d3q19_download_macro(h, s.a.rho, nullptr, nullptr, nullptr);
}

extern "C" int d3q19_shim_sync_f_to_host(d3q19_handle *h) {
    Shim &s = h->shim;
    if (!s.bound) return fail("d3q19_shim_*: call d3q19_shim_bind first");
    if (s.f_dev_valid && !s.f_host_valid) {
        RK_(d3q19_download_f(h, s.a.f));
        s.f_host_valid = true;
    }
    return 0;
}

extern "C" int d3q19_shim_sync_f_to_device(d3q19_handle *h) {
    Shim &s = h->shim;
    if (!s.bound) return fail("d3q19_shim_*: call d3q19_shim_bind first");
    s.f_dev_valid = false;
    s.f_host_valid = true;
    s.macro_dev_valid = false;
    s.u_dev_loaded = false;
    return 0;
}
