// liblbm_b200.so -- context management and the C ABI declared in include/lbm_b200.h.
// Kernels live in kernels_inst.cu (one instance per <lattice, arithmetic mode>).
#include <cuda_runtime.h>
#include <dlfcn.h>
#include <unistd.h>
#include <nccl.h>  // types only; the library is dlopen'ed when world > 1

#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <algorithm>
#include <string>
#include <cuda.h>
#include <thread>
#include <vector>

#include <type_traits>
#include "common.h"

namespace lbm {
#define LBM_DECL(a, b) const Ops *get_ops_##a##_##b();
#define LBM_DECL_BOTH(a) LBM_DECL(a, 0) LBM_DECL(a, 1)
LBM_DECL_BOTH(LBM_D2Q4) LBM_DECL_BOTH(LBM_D2Q5) LBM_DECL_BOTH(LBM_D2Q9) LBM_DECL_BOTH(LBM_D2Q13)
LBM_DECL_BOTH(LBM_D2Q17) LBM_DECL_BOTH(LBM_D2Q21) LBM_DECL_BOTH(LBM_D2Q37)

const Ops *get_ops(int lattice, int arith) {
#define LBM_CASE(a) case a: return arith ? get_ops_##a##_1() : get_ops_##a##_0();
    switch (lattice) {
        LBM_CASE(LBM_D2Q4) LBM_CASE(LBM_D2Q5) LBM_CASE(LBM_D2Q9) LBM_CASE(LBM_D2Q13)
        LBM_CASE(LBM_D2Q17) LBM_CASE(LBM_D2Q21) LBM_CASE(LBM_D2Q37)
    default: return nullptr;
    }
}
}  // namespace lbm

using namespace lbm;

// ----------------------------------------------------------------------------------------------
// errors
// ----------------------------------------------------------------------------------------------
static thread_local std::string g_err;

static int fail(int code, const char *fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    g_err = buf;
    return code;
}

#define CU(call)                                                                                         \
    do {                                                                                                 \
        cudaError_t e_ = (call);                                                                         \
        if (e_ != cudaSuccess) return fail(LBM_ERR_CUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__); \
    } while (0)

// ----------------------------------------------------------------------------------------------
// NCCL, loaded lazily
// ----------------------------------------------------------------------------------------------
struct NcclApi {
    void *handle = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId *) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*Send)(const void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Recv)(void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*AllReduce)(const void *, void *, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    const char *(*GetErrorString)(ncclResult_t) = nullptr;
};
static NcclApi g_nccl;

static std::mutex g_nccl_mutex;  // contexts of one process may be created from several host threads

static int load_nccl() {
    std::lock_guard<std::mutex> lock(g_nccl_mutex);
    if (g_nccl.handle) return 0;
    const char *names[] = {getenv("LBM_NCCL_LIB"), "libnccl.so.2", "libnccl.so"};
    void *h = nullptr;
    for (const char *n : names) {
        if (!n) continue;
        h = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
        if (h) break;
    }
    if (!h) return fail(LBM_ERR_NCCL, "cannot dlopen libnccl.so.2: %s", dlerror());
#define SYM(field, name)                                                     \
    *(void **)(&g_nccl.field) = dlsym(h, name);                              \
    if (!g_nccl.field) return fail(LBM_ERR_NCCL, "missing NCCL symbol %s", name);
    SYM(GetUniqueId, "ncclGetUniqueId") SYM(CommInitRank, "ncclCommInitRank") SYM(CommDestroy, "ncclCommDestroy")
    SYM(Send, "ncclSend") SYM(Recv, "ncclRecv") SYM(GroupStart, "ncclGroupStart") SYM(GroupEnd, "ncclGroupEnd")
    SYM(GetErrorString, "ncclGetErrorString") SYM(AllReduce, "ncclAllReduce")
#undef SYM
    g_nccl.handle = h;
    return 0;
}

#define NC(call)                                                                                    \
    do {                                                                                            \
        ncclResult_t r_ = (call);                                                                   \
        if (r_ != ncclSuccess) return fail(LBM_ERR_NCCL, "%s failed: %s", #call, g_nccl.GetErrorString(r_)); \
    } while (0)

// ----------------------------------------------------------------------------------------------
// context
// ----------------------------------------------------------------------------------------------
enum { ST_STREAM = 0, ST_COLLIDED = 1 };

// What a rank tells its two neighbours about its exported allocation, and the neighbour's view of it.
struct PeerInfo {
    cudaIpcMemHandle_t handle;  // of the cudaMalloc allocation that contains the arena
    unsigned long long offset;  // arena - base of that allocation
    unsigned long long raw;     // arena address in the owner's process (peers living in the same process)
    unsigned long long buf_bytes;
    long long plane;
    int nyl, pid, device, rank;
};
struct PeerMap {
    PeerInfo info;
    void *ipc_base = nullptr;  // from cudaIpcOpenMemHandle (nullptr: same process or shared with the other neighbour)
    char *arena = nullptr;     // neighbour's arena in this process' address space
};
static const size_t ARENA_HDR = 4096;
static const unsigned long long P2P_MAGIC_VALUE = 0x4c424d5032500000ULL;  // "LBMP2P"

struct lbm_ctx {
    lbm_desc desc;
    LatticeInfo li;
    const Ops *ops = nullptr;
    int nyl = 0, y0 = 0;
    int gx = 0, gy = 0;
    long long pitch = 0, plane = 0;
    size_t elt = 8;
    void *buf[2] = {nullptr, nullptr};
    int cur = 0;         // buffer holding the newest state
    int state = ST_STREAM;
    bool have_coll = false;  // buf[1-cur] holds f_collision matching f_stream in buf[cur]
    bool resume_ok = false;  // ... and its ghosts/halos are valid, so the next step may pull from it
    cudaStream_t stream = nullptr, comm_stream = nullptr, bstream = nullptr;  // bstream: boundary-row kernels
    cudaEvent_t ev_b = nullptr, ev_c = nullptr, ev_t0 = nullptr, ev_t1 = nullptr, ev_u0 = nullptr, ev_u1 = nullptr, ev_fork = nullptr;
    bool comm_pending = false;
    // force
    int force_mode = 0;
    double fx = 0, fy = 0;
    void *field = nullptr;
    void *sep_fx = nullptr, *sep_fy = nullptr;
    long long sep_t0 = 0;
    int sep_n = 0;
    // diagnostics
    double *partials = nullptr, *red_out = nullptr, *u_old = nullptr, *rho_old = nullptr;
    int red_blocks = 8192;
    // multi-GPU
    ncclComm_t comm = nullptr;
    int up = 0, down = 0;
    // peer-memory halo exchange (world > 1): one exported allocation [flag block | buf 0 | buf 1] per rank
    void *arena = nullptr;
    size_t buf_bytes = 0;
    PeerMap peer_up, peer_dn;
    bool p2p_on = false;       // neighbours' arenas are mapped and every rank agreed to use them
    int opt_p2p = 1;           // lbm_set_option("p2p", 0) falls back to NCCL send/recv (same on all ranks)
    unsigned long long epoch = 0, batch = 0;
    bool p2p_in_batch = false; // a P2P launch ran in the current lbm_step batch
    unsigned long long p2p_failed = 0;  // sticky: a device-side wait of the protocol gave up (epoch/token it waited for)
    // CUDA graphs of GRAPH_STEPS fused steps (launch-bound small slabs): one per source buffer, rebuilt when
    // anything baked into the kernel parameters changes (force data, options)
    cudaGraphExec_t graph[2] = {nullptr, nullptr};
    long long graph_launches[2] = {0, 0};  // kernels inside each graph
    int opt_graph = 1;
    bool capturing = false;
    unsigned long long cap_rel = 0;        // relative epoch of the launch being captured
    long long launches = 0;
    bool timed = false;
    int opt_variant = 0;
    int opt_overlap = 3;  // y-slabs: 0 whole slab after the exchange, 1 boundary + interior launches, 2 merged launch, 3 automatic
    // persistent multi-step kernel (persist.cuh): 0 = off, 1 = whenever possible, 2 = automatic (= off: measured slower, see persist_ok)
    int opt_persistent = 2;
    int pg_ctas[2] = {-1, -1}, pg_threads[2] = {0, 0};  // co-resident grid of the plain / peer-memory variant (-1: not queried)
    unsigned long long *pdone = nullptr;                // [error | edge_count[2] | pad | done[ctas]]
    int pdone_ctas = 0;
    bool persist_used = false;
    unsigned long long persist_failed = 0;
    // lbm_reduce_errors scratch: [separable tables | partials | out], kept between calls
    double *err_dev = nullptr;
    size_t err_doubles = 0;
    std::vector<double> err_tab, scratch_row;  // host copy of the tables on the device, per (field, term) slot
    unsigned err_tab_valid = 0;                // bit slot: err_tab[slot] is what the device holds
    // TMA-staged fused step (tma.cuh): the two population buffers as 3-D tensor maps (pitch, rows incl. ghosts, Q)
    alignas(64) CUtensorMap tmap[2];
    bool tma_ok = false;
    int opt_store = 0;      // tuning builds: cache operator of the population stores
    int opt_prefetch = -1;  // L2-prefetch distance (rows) of the fused pull; 0 = off, -1 = automatic (prefetch_rows)
    int opt_tma = 2;      // 0 = never, 1 = wherever available, 2 = automatic (tma_auto)
    int opt_tma_cfg = 0;  // tuning: 100 * (CTA width / 128) + 10 * stages + CTAs per SM; 0 = default
    // page-locked chunk buffers for copies from / to pageable host arrays (HostPipe)
    char *pipe_pin[4] = {nullptr, nullptr, nullptr, nullptr};
    cudaEvent_t pipe_ev[4] = {nullptr, nullptr, nullptr, nullptr};
    size_t pipe_bytes = 0;
    // lbm_moments output fields on the device, kept between calls
    double *mom_dev = nullptr;
    size_t mom_bytes = 0;
    // asynchronous snapshots (lbm_snapshot_begin / _end)
    cudaStream_t copy_stream = nullptr;
    cudaEvent_t ev_snap = nullptr, ev_snap_done = nullptr;
    double *snap_dev = nullptr, *snap_host = nullptr, *snap_user = nullptr;
    size_t snap_dev_bytes = 0, snap_host_bytes = 0;
    bool snap_pending = false, snap_direct = false;
    // reusable device staging for the Float32 import/export conversions (grown on demand, freed by lbm_destroy)
    double *stage = nullptr;
    size_t stage_bytes = 0;
};

static int need_stage(lbm_ctx *c, size_t bytes) {
    if (c->stage_bytes >= bytes) return 0;
    if (c->stage) { cudaStreamSynchronize(c->stream); cudaFree(c->stage); c->stage = nullptr; c->stage_bytes = 0; }
    cudaError_t e = cudaMalloc(&c->stage, bytes);
    if (e != cudaSuccess) { cudaGetLastError(); return fail(LBM_ERR_NOMEM, "cudaMalloc(%zu bytes of staging): %s", bytes, cudaGetErrorString(e)); }
    c->stage_bytes = bytes;
    return 0;
}

template <typename T>
static T *origin(const lbm_ctx *c, int b) {
    return reinterpret_cast<T *>(c->buf[b]) + (long long)c->gy * c->pitch + c->gx;
}

// Collision constants derived from the relaxation times, already converted to T (KParams<T> and BatchConsts<T> carry
// the same members):
//   SRT: c[0] = 1 - 1/tau, c[1] = 1/tau, shift = tau
//   TRT: c[0] = -(1/tau_s), c[1] = 1/tau_a, shift = tau_a
//   MRT: c[2n] = 1 - 1/tau_n, c[2n+1] = 1/tau_n (n = 2..N), kn[n] = css^n / n!, shift = tau_2
template <typename T, class P>
static void fill_collision_consts(P &p, int collision, const double *tau, int ntau, const LatticeInfo &li) {
    const double css = li.css;
    switch (collision) {
    case LBM_SRT:
    case LBM_ITERATIVE_INIT:
        p.c[0] = (T)(1 - 1 / tau[0]); p.c[1] = (T)(1 / tau[0]); p.shift = (T)tau[0];
        break;
    case LBM_TRT:
        p.c[0] = (T)(-(1 / tau[0])); p.c[1] = (T)(1 / tau[1]); p.shift = (T)tau[1];
        break;
    default: {
        static const double fact[5] = {1, 1, 2, 6, 24};
        for (int n = 2; n <= li.N; ++n) {
            const double tn = tau[n - 1];
            p.c[2 * n] = (T)(1 - 1 / tn);
            p.c[2 * n + 1] = (T)(1 / tn);
            p.kn[n] = (T)(std::pow(css, (double)n) / fact[n]);
            p.mrt_skip[n] = (tn == 1.0);
        }
        p.shift = (T)(ntau >= 2 ? tau[1] : 0.0);
        break;
    }
    }
}

static int prefetch_rows(const lbm_ctx *c);

template <typename T>
static KParams<T> make_params(const lbm_ctx *c, int src, int dst) {
    KParams<T> p;
    memset(&p, 0, sizeof(p));
    p.src = origin<T>(c, src);
    p.dst = origin<T>(c, dst);
    p.aux = nullptr;
    p.pitch = c->pitch;
    p.plane = c->plane;
    for (int i = 0; i < c->li.Q; ++i) {
        p.srcn[i] = p.src + (long long)i * c->plane;
        p.srcp[i] = p.srcn[i] - (long long)c->li.cy[i] * c->pitch - c->li.cx[i];
        p.dstp[i] = p.dst + (long long)i * c->plane;
    }
    p.nx = c->desc.nx;
    p.nyl = c->nyl;
    p.y0g = c->y0;
    p.nyg = c->desc.ny;
    p.row_a0 = 0; p.row_an = c->nyl; p.row_b0 = 0; p.row_bn = 1 << 30; p.row_c0 = 0; p.nrows = c->nyl;
    p.pf_rows = prefetch_rows(c);
    p.st_mode = c->opt_store;
    p.p2p_rows = 1 << 30;  // P2P launches: every CTA takes part unless the caller narrows it to the edge rows
    p.wrap_y = c->desc.world == 1;
    fill_collision_consts<T>(p, c->desc.collision, c->desc.tau, c->desc.ntau, c->li);
    p.force_mode = c->force_mode;
    p.fx = (T)c->fx; p.fy = (T)c->fy;
    p.field = (const T *)c->field;
    p.sep_fx = (const T *)c->sep_fx; p.sep_fy = (const T *)c->sep_fy;
    p.sep_t0 = c->sep_t0;
    p.nbc = c->desc.n_bcs;
    for (int b = 0; b < p.nbc; ++b) {
        const lbm_bc &s = c->desc.bcs[b];
        BCd &d = p.bc[b];
        d.kind = s.kind; d.dir = s.direction;
        d.x0 = s.x0; d.x1 = s.x1; d.y0 = s.y0; d.y1 = s.y1;
        d.ax = s.rho * s.u[0]; d.ay = s.rho * s.u[1];  // equilibrium_coefficient(Val{1}) hermite.jl:41-43
        p.bc_sides |= 1 << s.direction;
    }
    return p;
}

static bool is64(const lbm_ctx *c) { return c->desc.dtype == LBM_F64; }

// ----------------------------------------------------------------------------------------------
// halo exchange (y-slabs): rows of populations moving up go to `up`, moving down to `down`
// ----------------------------------------------------------------------------------------------
static int exchange_halos(lbm_ctx *c, int b, cudaStream_t s) {
    if (c->desc.world == 1) return 0;
    const size_t es = c->elt;
    char *base = (char *)c->buf[b];
    auto row = [&](int i, int y) {  // address of (plane i, local row y, x = -gx)
        return base + ((size_t)i * c->plane + (size_t)(y + c->gy) * c->pitch) * es;
    };
    const ncclDataType_t dt = is64(c) ? ncclDouble : ncclFloat;
    NC(g_nccl.GroupStart());
    // sends: first everything going up, then everything going down (same order on every rank,
    // which keeps send/recv matching correct when up == down, i.e. world == 2)
    for (int i = 0; i < c->li.Q; ++i) {
        const int cy = c->li.cy[i];
        if (cy > 0) NC(g_nccl.Send(row(i, c->nyl - cy), (size_t)cy * c->pitch, dt, c->up, c->comm, s));
    }
    for (int i = 0; i < c->li.Q; ++i) {
        const int cy = c->li.cy[i];
        if (cy < 0) NC(g_nccl.Send(row(i, 0), (size_t)(-cy) * c->pitch, dt, c->down, c->comm, s));
    }
    // receives: what `down` sent up lands in my bottom ghost rows, what `up` sent down in my top ghost rows
    for (int i = 0; i < c->li.Q; ++i) {
        const int cy = c->li.cy[i];
        if (cy > 0) NC(g_nccl.Recv(row(i, -cy), (size_t)cy * c->pitch, dt, c->down, c->comm, s));
    }
    for (int i = 0; i < c->li.Q; ++i) {
        const int cy = c->li.cy[i];
        if (cy < 0) NC(g_nccl.Recv(row(i, c->nyl), (size_t)(-cy) * c->pitch, dt, c->up, c->comm, s));
    }
    NC(g_nccl.GroupEnd());
    c->launches += 1;
    return 0;
}

// main stream waits until the last posted exchange has landed
static int wait_comm(lbm_ctx *c) {
    if (c->comm_pending) {
        CU(cudaStreamWaitEvent(c->stream, c->ev_c, 0));
        c->comm_pending = false;
    }
    return 0;
}

static int post_exchange(lbm_ctx *c, int b) {
    if (c->desc.world == 1) return 0;
    CU(cudaEventRecord(c->ev_b, c->stream));
    CU(cudaStreamWaitEvent(c->comm_stream, c->ev_b, 0));
    int rc = exchange_halos(c, b, c->comm_stream);
    if (rc) return rc;
    CU(cudaEventRecord(c->ev_c, c->comm_stream));
    c->comm_pending = true;
    return 0;
}

// ----------------------------------------------------------------------------------------------
// peer-memory halo exchange: mapping the neighbours' buffers (cudaIpc, or plain peer access inside one process)
// ----------------------------------------------------------------------------------------------
static unsigned long long *my_flags(const lbm_ctx *c) { return reinterpret_cast<unsigned long long *>(c->arena); }
static unsigned long long *peer_flags(const PeerMap &m) { return reinterpret_cast<unsigned long long *>(m.arena); }

static void p2p_unmap(lbm_ctx *c) {
    for (PeerMap *m : {&c->peer_up, &c->peer_dn}) {
        if (m->ipc_base) cudaIpcCloseMemHandle(m->ipc_base);
        m->ipc_base = nullptr;
        m->arena = nullptr;
    }
    c->p2p_on = false;
}

static bool p2p_map_one(lbm_ctx *c, PeerMap &m) {
    if (m.info.pid == (int)getpid()) {
        if (m.info.device != c->desc.device) {
            int can = 0;
            if (cudaDeviceCanAccessPeer(&can, c->desc.device, m.info.device) != cudaSuccess || !can) return false;
            cudaError_t e = cudaDeviceEnablePeerAccess(m.info.device, 0);
            if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) { cudaGetLastError(); return false; }
            cudaGetLastError();
        }
        m.arena = reinterpret_cast<char *>(m.info.raw);
    } else {
        void *base = nullptr;
        if (cudaIpcOpenMemHandle(&base, m.info.handle, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) { cudaGetLastError(); return false; }
        m.ipc_base = base;
        m.arena = reinterpret_cast<char *>(base) + m.info.offset;
    }
    // self-check through the mapping: the owner wrote MAGIC ^ rank into its flag block
    unsigned long long v = 0;
    if (cudaMemcpy(&v, m.arena + P2P_MAGIC * sizeof(unsigned long long), sizeof(v), cudaMemcpyDefault) != cudaSuccess) { cudaGetLastError(); return false; }
    return v == (P2P_MAGIC_VALUE ^ (unsigned long long)m.info.rank);
}

// Collective over all ranks of the communicator (called from lbm_create when world > 1).
static int p2p_connect(lbm_ctx *c) {
    PeerInfo mine;
    memset(&mine, 0, sizeof(mine));
    int local_ok = 1;
    if (const char *e = getenv("LBM_P2P")) if (atoi(e) == 0) local_ok = 0;
    // base of the cudaMalloc allocation the arena lives in (small arenas may be sub-allocated)
    unsigned long long base = (unsigned long long)c->arena;
    {
        typedef int (*GetRange)(unsigned long long *, size_t *, unsigned long long);
        void *fn = nullptr;
        cudaDriverEntryPointQueryResult qr;
        if (cudaGetDriverEntryPoint("cuMemGetAddressRange", &fn, cudaEnableDefault, &qr) == cudaSuccess && fn && qr == cudaDriverEntryPointSuccess) {
            unsigned long long b = 0; size_t sz = 0;
            if (((GetRange)fn)(&b, &sz, (unsigned long long)c->arena) == 0 && b) base = b;
        }
        cudaGetLastError();
    }
    if (cudaIpcGetMemHandle(&mine.handle, (void *)base) != cudaSuccess) { cudaGetLastError(); local_ok = 0; }
    mine.offset = (unsigned long long)c->arena - base;
    mine.raw = (unsigned long long)c->arena;
    mine.buf_bytes = c->buf_bytes;
    mine.plane = c->plane;
    mine.nyl = c->nyl; mine.pid = (int)getpid(); mine.device = c->desc.device; mine.rank = c->desc.rank;
    const unsigned long long magic = P2P_MAGIC_VALUE ^ (unsigned long long)c->desc.rank;
    CU(cudaMemcpy(my_flags(c) + P2P_MAGIC, &magic, sizeof(magic), cudaMemcpyHostToDevice));

    // ship the descriptors to both neighbours with the communicator we already have
    char *dev = nullptr;
    CU(cudaMalloc(&dev, 3 * sizeof(PeerInfo) + 16));
    cudaError_t e = cudaMemcpy(dev, &mine, sizeof(mine), cudaMemcpyHostToDevice);
    ncclResult_t r = ncclSuccess;
    if (e == cudaSuccess) {
        g_nccl.GroupStart();
        ncclResult_t r1 = g_nccl.Send(dev, sizeof(PeerInfo), ncclChar, c->up, c->comm, c->stream);
        ncclResult_t r2 = g_nccl.Send(dev, sizeof(PeerInfo), ncclChar, c->down, c->comm, c->stream);
        ncclResult_t r3 = g_nccl.Recv(dev + sizeof(PeerInfo), sizeof(PeerInfo), ncclChar, c->down, c->comm, c->stream);
        ncclResult_t r4 = g_nccl.Recv(dev + 2 * sizeof(PeerInfo), sizeof(PeerInfo), ncclChar, c->up, c->comm, c->stream);
        ncclResult_t r5 = g_nccl.GroupEnd();
        for (ncclResult_t x : {r1, r2, r3, r4, r5}) if (x != ncclSuccess) r = x;
        if (r == ncclSuccess) e = cudaStreamSynchronize(c->stream);
    }
    if (e == cudaSuccess && r == ncclSuccess) {
        e = cudaMemcpy(&c->peer_dn.info, dev + sizeof(PeerInfo), sizeof(PeerInfo), cudaMemcpyDeviceToHost);
        if (e == cudaSuccess) e = cudaMemcpy(&c->peer_up.info, dev + 2 * sizeof(PeerInfo), sizeof(PeerInfo), cudaMemcpyDeviceToHost);
    }
    if (e != cudaSuccess || r != ncclSuccess) {
        cudaFree(dev);
        if (r != ncclSuccess) return fail(LBM_ERR_NCCL, "peer descriptor exchange: %s", g_nccl.GetErrorString(r));
        return fail(LBM_ERR_CUDA, "peer descriptor exchange: %s", cudaGetErrorString(e));
    }
    if (local_ok) {
        local_ok = p2p_map_one(c, c->peer_up) ? 1 : 0;
        if (local_ok) {
            if (c->up == c->down) { c->peer_dn.arena = c->peer_up.arena; c->peer_dn.ipc_base = nullptr; }
            else local_ok = p2p_map_one(c, c->peer_dn) ? 1 : 0;
        }
    }
    // every rank must take the same path: min over ranks
    int *flag = reinterpret_cast<int *>(dev + 3 * sizeof(PeerInfo));
    e = cudaMemcpy(flag, &local_ok, sizeof(int), cudaMemcpyHostToDevice);
    if (e == cudaSuccess) {
        r = g_nccl.AllReduce(flag, flag, 1, ncclInt, ncclMin, c->comm, c->stream);
        if (r == ncclSuccess) e = cudaStreamSynchronize(c->stream);
    }
    int all_ok = 0;
    if (e == cudaSuccess && r == ncclSuccess) e = cudaMemcpy(&all_ok, flag, sizeof(int), cudaMemcpyDeviceToHost);
    cudaFree(dev);
    if (r != ncclSuccess) return fail(LBM_ERR_NCCL, "peer agreement: %s", g_nccl.GetErrorString(r));
    if (e != cudaSuccess) return fail(LBM_ERR_CUDA, "peer agreement: %s", cudaGetErrorString(e));
    if (all_ok) c->p2p_on = true;
    else p2p_unmap(c);
    return 0;
}

// after a stream synchronisation: did a device-side wait of the protocol give up?
// Called by every entry point that synchronises the stream and hands data back; the failure is sticky (lbm_step
// refuses to continue from a state computed with stale ghost rows).
static int p2p_check(lbm_ctx *c) {
    if (c->persist_used) {  // did a wait inside a persistent launch give up?
        if (!c->persist_failed) {
            unsigned long long v = 0;
            CU(cudaMemcpyAsync(&v, c->pdone, sizeof(v), cudaMemcpyDeviceToHost, c->stream));
            CU(cudaStreamSynchronize(c->stream));
            c->persist_failed = v;
        }
        if (c->persist_failed)
            return fail(LBM_ERR_STATE, "a wait inside the persistent step kernel timed out (step %llu of its launch); the populations "
                                       "of this context are invalid", c->persist_failed - 1);
    }
    if (!c->p2p_on || (c->epoch == 0 && c->batch == 0)) return 0;
    if (!c->p2p_failed) {
        unsigned long long v = 0;
        CU(cudaMemcpyAsync(&v, my_flags(c) + P2P_TIMEOUT, sizeof(v), cudaMemcpyDeviceToHost, c->stream));
        CU(cudaStreamSynchronize(c->stream));
        c->p2p_failed = v;
    }
    if (c->p2p_failed)
        return fail(LBM_ERR_STATE, "peer-memory halo exchange timed out waiting for a neighbour (epoch/token %llu); the "
                                   "populations of this context are invalid", c->p2p_failed);
    return 0;
}

// L2-prefetch distance of the fused pull kernel (rows ahead of the row a CTA is working on; one lane per 128-byte line
// issues prefetch.global.L2 for every population).  Measured on B200 (profiles/r02/prefetch_sweep_v2.jsonl, fraction of the
// measured copy peak, off -> best): Float64 D2Q9 TRT 0.982 -> 1.019 (16 rows), D2Q9 MRT 0.976 -> 1.017, D2Q13 0.994 -> 1.004
// (8), D2Q17 TRT 0.951 -> 0.964 (16), D2Q17 MRT 0.909 -> 0.939 (24), D2Q21 0.951 -> 0.962 (12), D2Q37 TRT 0.936 -> 0.943 (4),
// D2Q37 MRT 0.855 -> 0.871 (16); distances of 64+ rows lose (the lines are evicted before use).  Float32 loses ~14 % at
// every distance (those kernels are issue-bound: Q more instructions per node cost more than the latency they hide).
// What matters is the distance in NODES (the sweep ran at 4096 / 2048 nodes per row): the same number of rows on a
// 32768-wide grid is 8x further ahead in time and the lines are evicted before use (32768^2 fell from 45.3 to 33.0 GLUPS
// with a fixed 16 rows), so the automatic distance is a node count converted to rows of this grid.
static int prefetch_rows(const lbm_ctx *c) {
    if (c->opt_prefetch >= 0) return c->opt_prefetch;
    if (c->desc.dtype != LBM_F64) return 0;
    const bool mrt = c->desc.collision == LBM_MRT;
    const int Q = c->li.Q;
    long long nodes;  // best distance of the sweep x nodes per row of the sweep
    if (Q <= 9) nodes = 16 * 4096;
    else if (Q <= 13) nodes = 8 * 4096;
    else if (Q <= 17) nodes = (mrt ? 24 : 16) * 2048;
    else if (Q <= 25) nodes = 12 * 2048;
    else nodes = (mrt ? 16 : 4) * 2048;
    const long long nx = c->desc.nx;
    long long rows = (nodes + nx / 2) / nx;
    if (rows < 1) rows = 1;
    // small slabs are launch-bound and only pay the extra instructions (512^2: 0.71 -> 0.67 with prefetch)
    if (rows * 4 > c->nyl || nx * c->nyl < (1LL << 21)) return 0;
    return (int)rows;
}

template <typename T>
static void fill_p2p(const lbm_ctx *c, KParams<T> &p, int dst) {
    const size_t org = ((size_t)c->gy * c->pitch + c->gx) * c->elt;
    p.peer_up = reinterpret_cast<T *>(c->peer_up.arena + ARENA_HDR + (size_t)dst * c->peer_up.info.buf_bytes + org);
    p.peer_dn = reinterpret_cast<T *>(c->peer_dn.arena + ARENA_HDR + (size_t)dst * c->peer_dn.info.buf_bytes + org);
    p.plane_up = c->peer_up.info.plane;
    p.plane_dn = c->peer_dn.info.plane;
    p.nyl_dn = c->peer_dn.info.nyl;
    p.flags = my_flags(c);
    p.flag_at_up = peer_flags(c->peer_up) + P2P_EPOCH_FROM_DOWN;
    p.flag_at_dn = peer_flags(c->peer_dn) + P2P_EPOCH_FROM_UP;
}

// ----------------------------------------------------------------------------------------------
// kernel sequencing
// ----------------------------------------------------------------------------------------------
// Use the TMA-staged kernel (tma.cuh) for this context's fused pull launches?  Only on request: measured slower than or
// equal to the register kernel on every <lattice, dtype> (tma.cuh header, profiles/r02/tma_sweep_v2.jsonl), so automatic
// mode (2) keeps it off.
static bool tma_wanted(const lbm_ctx *c) {
    return c->opt_tma == 1 && c->desc.collision != LBM_ITERATIVE_INIT;
}

template <typename T>
static void run_step(lbm_ctx *c, bool pull, const KParams<T> &p, long long step, cudaStream_t s = nullptr) {
    if (!s) s = c->stream;
    c->launches += 1;
    if (pull && c->tma_ok && tma_wanted(c)) {
        const int b = (const void *)p.src == (const void *)origin<T>(c, 0) ? 0 : 1;
        int rc;
        if (std::is_same<T, double>::value) rc = c->ops->step_tma64(c->desc.collision, reinterpret_cast<const KParams<double> &>(p), &c->tmap[b], step, c->gx, c->gy, c->opt_tma_cfg, s);
        else rc = c->ops->step_tma32(c->desc.collision, reinterpret_cast<const KParams<float> &>(p), &c->tmap[b], step, c->gx, c->gy, c->opt_tma_cfg, s);
        if (rc == 0) return;
    }
    if (std::is_same<T, double>::value) c->ops->step64(c->desc.collision, pull, reinterpret_cast<const KParams<double> &>(p), step, c->opt_variant, s);
    else c->ops->step32(c->desc.collision, pull, reinterpret_cast<const KParams<float> &>(p), step, c->opt_variant, s);
}

// fused launch that also pushes its boundary rows into the neighbours' ghost rows
template <typename T>
static void run_step_p2p(lbm_ctx *c, KParams<T> &p, int dst, long long step, cudaStream_t s) {
    fill_p2p<T>(c, p, dst);
    if (c->capturing) {  // replayed launches: epoch relative to the base the host sets before each replay
        p.epoch = ++c->cap_rel;
        p.epoch_base = my_flags(c) + P2P_EPOCH_BASE;
    } else {
        p.epoch = ++c->epoch;
        c->p2p_in_batch = true;
    }
    if (std::is_same<T, double>::value) c->ops->step64_p2p(c->desc.collision, reinterpret_cast<const KParams<double> &>(p), step, c->opt_variant, s);
    else c->ops->step32_p2p(c->desc.collision, reinterpret_cast<const KParams<float> &>(p), step, c->opt_variant, s);
    c->launches += 1;
}

// collide buf[cur] (f_stream) -> buf[1-cur] (f_collision, with ghosts + halos)
template <typename T>
static int do_collide(lbm_ctx *c, long long step) {
    KParams<T> p = make_params<T>(c, c->cur, 1 - c->cur);
    run_step<T>(c, false, p, step);
    CU(cudaGetLastError());
    return post_exchange(c, 1 - c->cur);
}

// one fused step: buf[src] holds post-collision populations of the previous step
// Peer-memory y-slabs: one merged launch per step, or a boundary-row launch next to an interior launch?  Measured on
// 2 x B200 (profiles/r02/r16_*.json): 1024 x 1024 per GPU 69.3 vs 65.1 GLUPS (the merged form saves a kernel node and the
// fork / join events: 30.3 vs 32.2 us per step), 1024 x 4096 per GPU equal, 4096 x 4096 per GPU 88.7 vs 89.9 GLUPS (there
// the edge rows running beside the interior on a high-priority stream are worth more).  Automatic: merged up to 4 Mi nodes.
static bool merged_launch(const lbm_ctx *c) {
    if (c->opt_overlap == 2) return true;
    return c->opt_overlap == 3 && (long long)c->desc.nx * c->nyl <= (1LL << 22);
}

template <typename T>
static int do_fused(lbm_ctx *c, int src, int dst, long long step) {
    KParams<T> p = make_params<T>(c, src, dst);
    const int H = c->li.H;
    if (c->desc.world == 1) {
        run_step<T>(c, true, p, step);
    } else if (!c->opt_overlap || c->nyl < 2 * H + 1) {
        int rc = wait_comm(c);
        if (rc) return rc;
        if (c->p2p_on && c->opt_p2p) {  // whole slab in one launch; it pushes its own boundary rows
            run_step_p2p<T>(c, p, dst, step, c->stream);
            CU(cudaGetLastError());
            return 0;
        }
        run_step<T>(c, true, p, step);
    } else if (c->p2p_on && c->opt_p2p && merged_launch(c)) {
        // Merged launch: ONE kernel per step.  Its first CTAs (block order = launch order) take the 2H edge rows: they wait
        // for the neighbours' previous epoch, compute, store locally and into the neighbours' ghost rows, and publish
        // this epoch as soon as the edge rows are done; the remaining CTAs do the interior rows, which read no ghost
        // row, and take no part in the protocol.  Saves a kernel node and the fork / join events of the two-launch form.
        if (c->comm_pending) { CU(cudaStreamWaitEvent(c->stream, c->ev_c, 0)); c->comm_pending = false; }
        KParams<T> pm = p;
        pm.row_a0 = 0; pm.row_an = H; pm.row_b0 = c->nyl - H; pm.row_bn = H; pm.row_c0 = H; pm.nrows = c->nyl;
        pm.p2p_rows = 2 * H;
        run_step_p2p<T>(c, pm, dst, step, c->stream);
        CU(cudaGetLastError());
        return 0;
    } else if (c->p2p_on && c->opt_p2p) {
        // Peer-memory exchange: the boundary launch itself writes its rows into the neighbours' ghost rows and
        // hand-shakes through flags in peer memory (kernels_inst.cu), concurrently with the interior launch:
        //   main   : [fork] interior(t) .................. [join ev_b]
        //   bstream: wait fork (+ a pending NCCL exchange of the batch's first state); boundary+push(t); record ev_b
        CU(cudaEventRecord(c->ev_fork, c->stream));
        CU(cudaStreamWaitEvent(c->bstream, c->ev_fork, 0));
        if (c->comm_pending) { CU(cudaStreamWaitEvent(c->bstream, c->ev_c, 0)); c->comm_pending = false; }
        KParams<T> pb = p;
        pb.row_a0 = 0; pb.row_an = H; pb.row_b0 = c->nyl - H; pb.nrows = 2 * H;
        run_step_p2p<T>(c, pb, dst, step, c->bstream);
        CU(cudaEventRecord(c->ev_b, c->bstream));
        KParams<T> pi = p;
        pi.row_a0 = H; pi.row_an = c->nyl - 2 * H; pi.nrows = pi.row_an;
        run_step<T>(c, true, pi, step);
        CU(cudaGetLastError());
        CU(cudaStreamWaitEvent(c->stream, c->ev_b, 0));  // join
        return 0;
    } else {
        // The 2H boundary rows are the only ones that read halos and the only ones the neighbours need.
        // They run on a high-priority side stream concurrently with the interior kernel:
        //   main   : [fork] interior(t) ........................ [join ev_b]
        //   bstream: wait fork, wait exchange(t-1); boundary(t); record ev_b
        //   comm   : wait ev_b; NCCL send/recv of the boundary rows (overlaps interior(t+1)); record ev_c
        CU(cudaEventRecord(c->ev_fork, c->stream));  // everything of step t-1 (incl. its boundary rows) is done
        CU(cudaStreamWaitEvent(c->bstream, c->ev_fork, 0));
        if (c->comm_pending) CU(cudaStreamWaitEvent(c->bstream, c->ev_c, 0));
        KParams<T> pb = p;
        pb.row_a0 = 0; pb.row_an = H; pb.row_b0 = c->nyl - H; pb.nrows = 2 * H;
        run_step<T>(c, true, pb, step, c->bstream);
        CU(cudaEventRecord(c->ev_b, c->bstream));
        KParams<T> pi = p;
        pi.row_a0 = H; pi.row_an = c->nyl - 2 * H; pi.nrows = pi.row_an;
        run_step<T>(c, true, pi, step);
        CU(cudaGetLastError());
        CU(cudaStreamWaitEvent(c->stream, c->ev_b, 0));  // join
        CU(cudaStreamWaitEvent(c->comm_stream, c->ev_b, 0));
        int rc = exchange_halos(c, dst, c->comm_stream);
        if (rc) return rc;
        CU(cudaEventRecord(c->ev_c, c->comm_stream));
        c->comm_pending = true;
        return 0;
    }
    CU(cudaGetLastError());
    return post_exchange(c, dst);
}

// f_stream := (stream + BCs)(f_collision) into the other buffer; afterwards cur = f_stream
template <typename T>
static int do_materialize(lbm_ctx *c) {
    if (c->state != ST_COLLIDED) return 0;
    int rc = wait_comm(c);
    if (rc) return rc;
    KParams<T> p = make_params<T>(c, c->cur, 1 - c->cur);
    if (std::is_same<T, double>::value) c->ops->stream64(reinterpret_cast<const KParams<double> &>(p), c->stream);
    else c->ops->stream32(reinterpret_cast<const KParams<float> &>(p), c->stream);
    c->launches += 1;
    CU(cudaGetLastError());
    c->cur = 1 - c->cur;
    c->state = ST_STREAM;
    c->have_coll = true;
    c->resume_ok = true;
    return 0;
}

static int materialize(lbm_ctx *c) { return is64(c) ? do_materialize<double>(c) : do_materialize<float>(c); }

// ----------------------------------------------------------------------------------------------
// CUDA graphs for the step loop
// ----------------------------------------------------------------------------------------------
static const int GRAPH_STEPS = 16;  // even: the buffer roles return to where they started

static void drop_graphs(lbm_ctx *c) {
    for (int b = 0; b < 2; ++b)
        if (c->graph[b]) { cudaGraphExecDestroy(c->graph[b]); c->graph[b] = nullptr; }
}

// Everything baked into a captured launch must be replay-invariant: no per-step force table, and the halo
// exchange either absent or done by the kernels themselves (peer memory).
static bool graph_ok(const lbm_ctx *c) {
    return c->opt_graph && c->force_mode != 3 && (c->desc.world == 1 || (c->p2p_on && c->opt_p2p));
}

// capture GRAPH_STEPS fused steps starting from post-collision populations in buf[src]
template <typename T>
static int build_graph(lbm_ctx *c, int src) {
    const long long l0 = c->launches;
    if (cudaStreamBeginCapture(c->stream, cudaStreamCaptureModeRelaxed) != cudaSuccess) { cudaGetLastError(); return -1; }
    c->capturing = true;
    c->cap_rel = 0;
    int rc = 0, s = src;
    for (int g = 0; g < GRAPH_STEPS && rc == 0; ++g, s = 1 - s) rc = do_fused<T>(c, s, 1 - s, 0);
    c->capturing = false;
    cudaGraph_t graph = nullptr;
    cudaError_t e = cudaStreamEndCapture(c->stream, &graph);
    const long long inside = c->launches - l0;
    c->launches = l0;
    if (rc != 0 || e != cudaSuccess || !graph) { if (graph) cudaGraphDestroy(graph); cudaGetLastError(); return -1; }
    e = cudaGraphInstantiate(&c->graph[src], graph, 0);
    cudaGraphDestroy(graph);
    if (e != cudaSuccess) { c->graph[src] = nullptr; cudaGetLastError(); return -1; }
    c->graph_launches[src] = inside;
    return 0;
}

// ----------------------------------------------------------------------------------------------
// persistent multi-step launch (persist.cuh)
// ----------------------------------------------------------------------------------------------
static const long long PERSIST_MIN_STEPS = 4;

// The population buffers as tensor maps for cp.async.bulk.tensor (driver entry point resolved at run time: the library
// does not link libcuda).  Failure just leaves the TMA path off.
static void make_tensor_maps(lbm_ctx *c) {
    typedef CUresult (*encode_t)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                 const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                 CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
    static encode_t encode = [] {
        void *fn = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess) { cudaGetLastError(); fn = nullptr; }
        return (encode_t)fn;
    }();
    c->tma_ok = false;
    if (!encode) return;
    const cuuint64_t rows = (cuuint64_t)(c->plane / c->pitch);
    if ((long long)rows * c->pitch != c->plane) return;  // padded planes (layout experiments): not a regular tensor
    const cuuint64_t dims[3] = {(cuuint64_t)c->pitch, rows, (cuuint64_t)c->li.Q};
    const cuuint64_t strides[2] = {(cuuint64_t)c->pitch * c->elt, (cuuint64_t)c->plane * c->elt};
    const cuuint32_t estr[3] = {1, 1, 1};
    // box = one row segment of one population: the 128 nodes of a tile plus 4 elements on either side (tma.cuh: the box
    // must start 16-byte aligned, the x shift of the pull happens in the shared-memory read)
    const cuuint32_t box[3] = {128 + 8, 1, 1};
    if (c->gx < 4 || (c->gx * c->elt) % 16 != 0) return;  // narrow grids (4 ghost columns) have no aligned box start
    for (int b = 0; b < 2; ++b) {
        CUresult r = encode(&c->tmap[b], c->elt == 8 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT64 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, c->buf[b],
                            dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                            CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) return;
    }
    c->tma_ok = true;
}

static bool persist_ok(lbm_ctx *c, long long nsteps) {
    if (!c->opt_persistent || nsteps < PERSIST_MIN_STEPS || c->desc.collision == LBM_ITERATIVE_INIT) return false;
    const bool p2p = c->desc.world > 1;
    if (p2p && !(c->p2p_on && c->opt_p2p && c->opt_overlap && c->nyl >= 2 * c->li.H + 1)) return false;
    // two co-resident grids that wait for each other cannot share a GPU
    if (p2p && (c->peer_up.info.device == c->desc.device || c->peer_dn.info.device == c->desc.device)) return false;
    const long long N = (long long)c->desc.nx * c->nyl;
    if (N >= (1LL << 32)) return false;
    // Automatic mode (2) never selects it.  Measured in round 2 (profiles/r02): on ONE GPU graph replays of the per-step
    // launches beat the persistent kernel at every size (128^2: 4.1 vs 9.5 us per step, 1024^2: 27.5 vs 32.2 us), and on
    // 2 x B200 y-slabs of 1024 x 1024 it ran at 32.1 GLUPS against 59.8 GLUPS for boundary + interior launches from graphs
    // (r9_C3q_n2_p1 / _p0): every step ends with threadfence -> release -> acquire -> L2-latency reload in every CTA,
    // which costs more than a kernel boundary inside a graph.  It stays available as option persistent = 1 (tests keep it
    // bit-identical to the oracle).
    if (c->opt_persistent == 2) return false;
    int &ctas = c->pg_ctas[p2p ? 1 : 0];
    if (ctas < 0) {
        if (is64(c)) c->ops->persist_grid64(c->desc.collision, p2p, &ctas, &c->pg_threads[p2p ? 1 : 0]);
        else c->ops->persist_grid32(c->desc.collision, p2p, &ctas, &c->pg_threads[p2p ? 1 : 0]);
    }
    return ctas > 0;
}

// m fused steps from post-collision populations in buf[cur], one launch
template <typename T>
static int do_persist(lbm_ctx *c, long long t0, long long m) {
    const bool p2p = c->desc.world > 1;
    const int ctas = c->pg_ctas[p2p ? 1 : 0], threads = c->pg_threads[p2p ? 1 : 0];
    int rc = wait_comm(c);
    if (rc) return rc;
    if (c->pdone_ctas < ctas) {
        if (c->pdone) { CU(cudaStreamSynchronize(c->stream)); cudaFree(c->pdone); c->pdone = nullptr; }
        CU(cudaMalloc(&c->pdone, (size_t)(ctas + 4) * sizeof(unsigned long long)));
        CU(cudaMemsetAsync(c->pdone, 0, (size_t)(ctas + 4) * sizeof(unsigned long long), c->stream));
        c->pdone_ctas = ctas;
    }
    CU(cudaMemsetAsync(c->pdone + 1, 0, (size_t)(ctas + 3) * sizeof(unsigned long long), c->stream));  // all but the sticky error word
    const long long N = (long long)c->desc.nx * c->nyl, H = c->li.H;
    PersistArgs a;
    memset(&a, 0, sizeof(a));
    a.nsteps = (int)m;
    a.step0 = t0;
    a.error = c->pdone; a.edge_count = c->pdone + 1; a.done = c->pdone + 4;
    a.epoch0 = c->epoch;
    a.npc = (N + ctas - 1) / ctas;
    a.nctas = (int)((N + a.npc - 1) / a.npc);
    a.n_bot = (int)((H * c->desc.nx - 1) / a.npc) + 1;
    a.n_top = a.nctas - (int)(((long long)(c->nyl - H) * c->desc.nx) / a.npc);
    KParams<T> pa = make_params<T>(c, c->cur, 1 - c->cur), pb = make_params<T>(c, 1 - c->cur, c->cur);
    if (p2p) {
        fill_p2p<T>(c, pa, 1 - c->cur);
        fill_p2p<T>(c, pb, c->cur);
        c->epoch += (unsigned long long)m;
        c->p2p_in_batch = true;
    }
    int lrc;
    if (std::is_same<T, double>::value)
        lrc = c->ops->persist64(c->desc.collision, p2p, reinterpret_cast<const KParams<double> &>(pa), reinterpret_cast<const KParams<double> &>(pb), a, ctas, threads, c->stream);
    else
        lrc = c->ops->persist32(c->desc.collision, p2p, reinterpret_cast<const KParams<float> &>(pa), reinterpret_cast<const KParams<float> &>(pb), a, ctas, threads, c->stream);
    if (lrc != 0) return fail(LBM_ERR_CUDA, "cooperative launch of the persistent step kernel failed: %s", cudaGetErrorString(cudaGetLastError()));
    c->launches += 1;
    c->persist_used = true;
    if (m & 1) c->cur = 1 - c->cur;
    return 0;
}

template <typename T>
static int do_steps(lbm_ctx *c, long long t0, long long n) {
    for (long long k = 0; k < n; ++k) {
        const long long t = t0 + k;
        int rc;
        if (c->state == ST_COLLIDED && n - k <= 0x7fffffffLL && persist_ok(c, n - k)) {
            rc = do_persist<T>(c, t, n - k);
            if (rc) return rc;
            break;
        }
        if (c->state == ST_COLLIDED && !c->comm_pending && n - k >= GRAPH_STEPS && graph_ok(c)) {
            const int src = c->cur;
            if (!c->graph[src] && build_graph<T>(c, src) != 0) {
                c->opt_graph = 0;  // capture not possible here: plain launches from now on
            } else {
                if (c->desc.world > 1) {
                    c->ops->p2p_set_base(my_flags(c), c->epoch, c->stream);
                    c->launches += 1;
                    c->epoch += GRAPH_STEPS;
                    c->p2p_in_batch = true;
                }
                CU(cudaGraphLaunch(c->graph[src], c->stream));
                c->launches += c->graph_launches[src];
                k += GRAPH_STEPS - 1;
                continue;
            }
        }
        if (c->state == ST_STREAM) {
            if (c->resume_ok && c->have_coll) {
                // buf[1-cur] = f_collision of the previous step with valid ghosts: pull from it
                rc = do_fused<T>(c, 1 - c->cur, c->cur, t);
            } else {
                rc = do_collide<T>(c, t);
                c->cur = 1 - c->cur;
            }
            c->state = ST_COLLIDED;
            c->have_coll = false;
            c->resume_ok = false;
        } else {
            rc = do_fused<T>(c, c->cur, 1 - c->cur, t);
            c->cur = 1 - c->cur;
        }
        if (rc) return rc;
    }
    return 0;
}

// ----------------------------------------------------------------------------------------------
// C ABI
// ----------------------------------------------------------------------------------------------
extern "C" {

int lbm_abi_version(void) { return LBM_ABI_VERSION; }
const char *lbm_last_error(void) { return g_err.c_str(); }

int lbm_lattice_info(int32_t lattice, int32_t *q, int32_t *cx, int32_t *cy, double *w, double *css,
                     int32_t *opposite, int32_t *eq_order, int32_t *hermite_order, int32_t *halo) {
    LatticeInfo li;
    if (!lattice_info(lattice, li)) return fail(LBM_ERR_INVALID, "unknown lattice id %d", lattice);
    if (q) *q = li.Q;
    for (int i = 0; i < LBM_MAX_Q; ++i) {
        if (cx) cx[i] = li.cx[i];
        if (cy) cy[i] = li.cy[i];
        if (w) w[i] = li.w[i];
        if (opposite) opposite[i] = li.opp[i];
    }
    if (css) *css = li.css;
    if (eq_order) *eq_order = li.eq_order;
    if (hermite_order) *hermite_order = li.N;
    if (halo) *halo = li.H;
    return 0;
}

int lbm_nccl_unique_id(uint8_t id[LBM_NCCL_ID_BYTES]) {
    int rc = load_nccl();
    if (rc) return rc;
    ncclUniqueId uid;
    static_assert(sizeof(ncclUniqueId) == LBM_NCCL_ID_BYTES, "ncclUniqueId size");
    NC(g_nccl.GetUniqueId(&uid));
    memcpy(id, &uid, sizeof(uid));
    return 0;
}

void lbm_destroy(lbm_ctx *c) {
    if (!c) return;
    cudaSetDevice(c->desc.device);
    if (c->stream) cudaStreamSynchronize(c->stream);
    if (c->comm_stream) cudaStreamSynchronize(c->comm_stream);
    if (c->bstream) cudaStreamSynchronize(c->bstream);
    if (c->snap_pending) lbm_snapshot_end(c);  // the caller's array receives its snapshot before the context goes away
    if (c->copy_stream) { cudaStreamSynchronize(c->copy_stream); cudaStreamDestroy(c->copy_stream); }
    if (c->snap_dev) cudaFree(c->snap_dev);
    if (c->mom_dev) cudaFree(c->mom_dev);
    for (int k = 0; k < 4; ++k) {
        if (c->pipe_pin[k]) cudaFreeHost(c->pipe_pin[k]);
        if (c->pipe_ev[k]) cudaEventDestroy(c->pipe_ev[k]);
    }
    if (c->snap_host) cudaFreeHost(c->snap_host);
    for (cudaEvent_t e : {c->ev_snap, c->ev_snap_done})
        if (e) cudaEventDestroy(e);
    drop_graphs(c);
    p2p_unmap(c);
    if (c->comm && g_nccl.CommDestroy) g_nccl.CommDestroy(c->comm);
    if (c->arena) { cudaFree(c->arena); c->buf[0] = c->buf[1] = nullptr; }
    for (void *p : {c->buf[0], c->buf[1], c->field, c->sep_fx, c->sep_fy, (void *)c->partials, (void *)c->red_out, (void *)c->u_old, (void *)c->rho_old, (void *)c->stage, (void *)c->err_dev, (void *)c->pdone})
        if (p) cudaFree(p);
    for (cudaEvent_t e : {c->ev_b, c->ev_c, c->ev_t0, c->ev_t1, c->ev_u0, c->ev_u1, c->ev_fork})
        if (e) cudaEventDestroy(e);
    if (c->stream) cudaStreamDestroy(c->stream);
    if (c->comm_stream) cudaStreamDestroy(c->comm_stream);
    if (c->bstream) cudaStreamDestroy(c->bstream);
    delete c;
}

// everything about a descriptor that does not depend on the decomposition (shared by lbm_create / lbm_batch_create)
static int validate_desc(const lbm_desc *d, LatticeInfo &li) {
    if (d->abi_version != LBM_ABI_VERSION) return fail(LBM_ERR_INVALID, "abi_version %d != %d", d->abi_version, LBM_ABI_VERSION);
    if (d->nx < 1 || d->ny < 1) return fail(LBM_ERR_INVALID, "grid %dx%d", d->nx, d->ny);
    if (!lattice_info(d->lattice, li)) return fail(LBM_ERR_INVALID, "unknown lattice id %d", d->lattice);
    if (d->dtype != LBM_F64 && d->dtype != LBM_F32) return fail(LBM_ERR_INVALID, "dtype %d", d->dtype);
    if (d->collision < LBM_SRT || d->collision > LBM_ITERATIVE_INIT) return fail(LBM_ERR_INVALID, "collision %d", d->collision);
    if (d->arith != LBM_ARITH_EXACT && d->arith != LBM_ARITH_FAST) return fail(LBM_ERR_INVALID, "arith %d", d->arith);
    const int need_tau = d->collision == LBM_TRT ? 2 : (d->collision == LBM_MRT ? li.N : 1);
    if (d->ntau < need_tau || d->ntau > LBM_MAX_TAU) return fail(LBM_ERR_INVALID, "ntau %d (need >= %d)", d->ntau, need_tau);
    for (int i = 0; i < d->ntau; ++i)
        if (!(d->tau[i] == d->tau[i]) || d->tau[i] == 0.0) return fail(LBM_ERR_INVALID, "tau[%d] = %g", i, d->tau[i]);
    if (d->n_bcs < 0 || d->n_bcs > LBM_MAX_BCS) return fail(LBM_ERR_INVALID, "n_bcs %d", d->n_bcs);
    for (int b = 0; b < d->n_bcs; ++b) {
        const lbm_bc &bc = d->bcs[b];
        if (bc.direction < LBM_NORTH || bc.direction > LBM_WEST) return fail(LBM_ERR_INVALID, "bc %d: direction %d", b, bc.direction);
        if (bc.kind == LBM_BC_MOVING_WALL) {
            // only apply!(::MovingWall{<:North}, ...) exists (moving_wall.jl:17)
            if (bc.direction != LBM_NORTH) return fail(LBM_ERR_UNSUPPORTED, "bc %d: MovingWall is only defined for North", b);
        } else if (bc.kind != LBM_BC_BOUNCE_BACK) {
            return fail(LBM_ERR_INVALID, "bc %d: kind %d", b, bc.kind);
        }
    }
    return 0;
}

int lbm_create(const lbm_desc *d, lbm_ctx **out) {
    if (!d || !out) return fail(LBM_ERR_INVALID, "null argument");
    *out = nullptr;
    LatticeInfo li;
    int vrc = validate_desc(d, li);
    if (vrc) return vrc;
    if (d->world < 1 || d->rank < 0 || d->rank >= d->world) return fail(LBM_ERR_INVALID, "rank %d / world %d", d->rank, d->world);
    const Ops *ops = get_ops(d->lattice, d->arith);
    if (!ops) return fail(LBM_ERR_UNSUPPORTED, "no kernels for lattice %d", d->lattice);
    if (d->dtype == LBM_F32 && !ops->step32) return fail(LBM_ERR_UNSUPPORTED, "Float32 kernels not built");

    lbm_ctx *c = new lbm_ctx();
    c->desc = *d;
    c->li = li;
    c->ops = ops;
    // y-slabs: rows split as evenly as possible, the first (ny % world) ranks get one extra
    const int base = d->ny / d->world, rem = d->ny % d->world;
    c->nyl = base + (d->rank < rem ? 1 : 0);
    c->y0 = d->rank * base + (d->rank < rem ? d->rank : rem);
    if (d->world > 1 && c->nyl < li.H) {
        const int nyl = c->nyl;
        delete c;
        return fail(LBM_ERR_INVALID, "slab of %d rows is thinner than the halo (%d)", nyl, li.H);
    }
    c->elt = d->dtype == LBM_F64 ? 8 : 4;
    const int align = (int)(128 / c->elt);
    c->gx = d->nx >= 128 ? align : 4;
    c->gy = li.H;
    c->pitch = ((long long)d->nx + 2 * c->gx + align - 1) / align * align;
    // layout experiments (DRAM channel mapping of the Q concurrently streamed planes)
    if (const char *e = getenv("LBM_PAD_PITCH")) c->pitch += (long long)atoi(e) * align;
    long long plane_rows = c->nyl + 2 * c->gy;
    c->plane = c->pitch * plane_rows;
    if (const char *e = getenv("LBM_PAD_PLANE")) c->plane += (long long)atoi(e) * align;
    if (c->plane >= (1LL << 31)) {
        const long long pl = c->plane;
        delete c;
        return fail(LBM_ERR_INVALID, "slab of %lld elements per population exceeds the 2^31 node-index limit", pl);
    }
    c->up = (d->rank + 1) % d->world;
    c->down = (d->rank + d->world - 1) % d->world;

    cudaError_t e = cudaSetDevice(d->device);
    if (e != cudaSuccess) { delete c; return fail(LBM_ERR_CUDA, "cudaSetDevice(%d): %s", d->device, cudaGetErrorString(e)); }
    int rc = 0;
    auto bail = [&](int code) { lbm_destroy(c); return code; };
    if (ops->init_constants() != 0) return bail(fail(LBM_ERR_CUDA, "constant upload failed: %s", cudaGetErrorString(cudaGetLastError())));
    const size_t bytes = ((size_t)li.Q * c->plane * c->elt + 255) / 256 * 256;
    c->buf_bytes = bytes;
    if (d->world > 1) {
        // one allocation [flag block | buf 0 | buf 1] that the neighbours map (peer-memory halo exchange)
        size_t total = ARENA_HDR + 2 * bytes;
        if (total < ((size_t)2 << 20)) total = (size_t)2 << 20;
        e = cudaMalloc(&c->arena, total);
        if (e != cudaSuccess) return bail(fail(LBM_ERR_NOMEM, "cudaMalloc(%zu bytes): %s", total, cudaGetErrorString(e)));
        cudaMemset(c->arena, 0, total);
        c->buf[0] = (char *)c->arena + ARENA_HDR;
        c->buf[1] = (char *)c->arena + ARENA_HDR + bytes;
    } else {
        for (int b = 0; b < 2; ++b) {
            e = cudaMalloc(&c->buf[b], bytes);
            if (e != cudaSuccess) return bail(fail(LBM_ERR_NOMEM, "cudaMalloc(%zu bytes): %s", bytes, cudaGetErrorString(e)));
            cudaMemset(c->buf[b], 0, bytes);
        }
    }
    make_tensor_maps(c);
    if (cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking) != cudaSuccess ||
        cudaStreamCreateWithFlags(&c->comm_stream, cudaStreamNonBlocking) != cudaSuccess ||
        cudaStreamCreateWithPriority(&c->bstream, cudaStreamNonBlocking, -1) != cudaSuccess ||
        cudaEventCreateWithFlags(&c->ev_fork, cudaEventDisableTiming) != cudaSuccess ||
        cudaEventCreateWithFlags(&c->ev_b, cudaEventDisableTiming) != cudaSuccess ||
        cudaEventCreateWithFlags(&c->ev_c, cudaEventDisableTiming) != cudaSuccess ||
        cudaEventCreate(&c->ev_t0) != cudaSuccess || cudaEventCreate(&c->ev_t1) != cudaSuccess ||
        cudaEventCreate(&c->ev_u0) != cudaSuccess || cudaEventCreate(&c->ev_u1) != cudaSuccess)
        return bail(fail(LBM_ERR_CUDA, "stream/event creation failed: %s", cudaGetErrorString(cudaGetLastError())));
    if (cudaMalloc(&c->partials, (size_t)c->red_blocks * 4 * sizeof(double)) != cudaSuccess ||
        cudaMalloc(&c->red_out, 4 * sizeof(double)) != cudaSuccess)
        return bail(fail(LBM_ERR_NOMEM, "reduction buffers"));
    if (d->world > 1) {
        rc = load_nccl();
        if (rc) return bail(rc);
        ncclUniqueId uid;
        memcpy(&uid, d->nccl_id, sizeof(uid));
        ncclResult_t r = g_nccl.CommInitRank(&c->comm, d->world, uid, d->rank);
        if (r != ncclSuccess) return bail(fail(LBM_ERR_NCCL, "ncclCommInitRank: %s", g_nccl.GetErrorString(r)));
        cudaDeviceSynchronize();
        rc = p2p_connect(c);
        if (rc) return bail(rc);
    }
    cudaDeviceSynchronize();
    *out = c;
    return 0;
}

int lbm_local_rows(const lbm_ctx *c, int32_t *y0, int32_t *ny_local) {
    if (!c) return fail(LBM_ERR_INVALID, "null context");
    if (y0) *y0 = c->y0;
    if (ny_local) *ny_local = c->nyl;
    return 0;
}

static int refresh_ghosts(lbm_ctx *c, int b) {
    if (is64(c)) { KParams<double> p = make_params<double>(c, b, b); c->ops->ghosts64(p, c->stream); }
    else { KParams<float> p = make_params<float>(c, b, b); c->ops->ghosts32(p, c->stream); }
    c->launches += 1;
    CU(cudaGetLastError());
    return 0;
}

static bool is_pinned(const void *p);

// memcpy split over a few host threads: one thread moves ~10 GB/s, PCIe 5 x16 takes 55 GB/s
static void par_memcpy(void *dst, const void *src, size_t bytes) {
    static const int nthreads = [] {
        const char *e = getenv("LBM_COPY_THREADS");
        int n = e ? atoi(e) : (int)std::thread::hardware_concurrency() / 2;
        return n < 1 ? 1 : (n > 8 ? 8 : n);
    }();
    if (bytes < (1u << 20) || nthreads == 1) { memcpy(dst, src, bytes); return; }
    const size_t slice = ((bytes + nthreads - 1) / nthreads + 63) & ~(size_t)63;
    std::thread th[8];
    int started = 0;
    size_t covered = slice < bytes ? slice : bytes;  // [0, covered) is this thread's share; helpers take what follows
    for (int t = 1; t < nthreads; ++t) {
        const size_t off = (size_t)t * slice;
        if (off >= bytes) break;
        const size_t n = bytes - off < slice ? bytes - off : slice;
        try {
            th[started] = std::thread([=] { memcpy((char *)dst + off, (const char *)src + off, n); });
        } catch (...) {  // no more threads to be had (never let an exception cross the C ABI): copy the rest here
            break;
        }
        ++started;
        covered = off + n;
    }
    memcpy(dst, src, slice < bytes ? slice : bytes);
    if (covered < bytes) memcpy((char *)dst + covered, (const char *)src + covered, bytes - covered);
    for (int t = 0; t < started; ++t) th[t].join();
}

// Copies between PAGEABLE host arrays (what a caller's plain Array / numpy array is) and pitched device memory, pipelined
// through four page-locked chunk buffers: while the DMA engine moves chunk j, host threads copy chunk j + 1 (upload) or
// chunk j - 1 (download) between the caller's array and the page-locked buffer.  cudaMemcpy from pageable memory does the
// same staging inside the driver on one thread (6-10 GB/s measured); this path is bound by PCIe instead.  Page-locked
// arrays (lbm_host_alloc) skip it.
struct HostPipe {
    static constexpr int K = 4;
    static constexpr size_t CHUNK = 8u << 20;
    lbm_ctx *c;
    struct Slot { char *host = nullptr; size_t bytes = 0; bool busy = false; } slot[K];
    int next = 0;

    explicit HostPipe(lbm_ctx *c_) : c(c_) {}
    int prepare(size_t rowbytes) {
        const size_t need = rowbytes > CHUNK ? rowbytes : CHUNK;
        if (c->pipe_bytes >= need) return 0;
        for (int k = 0; k < K; ++k) {
            if (c->pipe_pin[k]) { cudaFreeHost(c->pipe_pin[k]); c->pipe_pin[k] = nullptr; }
            cudaError_t e = cudaHostAlloc((void **)&c->pipe_pin[k], need, cudaHostAllocDefault);
            if (e != cudaSuccess) { cudaGetLastError(); c->pipe_bytes = 0; return fail(LBM_ERR_NOMEM, "cudaHostAlloc(%zu bytes of copy staging): %s", need, cudaGetErrorString(e)); }
            if (!c->pipe_ev[k]) CU(cudaEventCreateWithFlags(&c->pipe_ev[k], cudaEventDisableTiming));
        }
        c->pipe_bytes = need;
        return 0;
    }
    int retire(int k) {  // wait for slot k's DMA; a download then lands in the caller's array
        if (!slot[k].busy) return 0;
        CU(cudaEventSynchronize(c->pipe_ev[k]));
        if (slot[k].host) par_memcpy(slot[k].host, c->pipe_pin[k], slot[k].bytes);
        slot[k].busy = false;
        return 0;
    }
    // `rows` rows of `rowbytes` bytes: host contiguous, device with pitch `dpitch`
    int put(char *dev, size_t dpitch, const char *host, size_t rowbytes, size_t rows) {
        int rc = prepare(rowbytes);
        if (rc) return rc;
        const size_t per = c->pipe_bytes / rowbytes;
        for (size_t r = 0; r < rows; r += per) {
            const size_t n = rows - r < per ? rows - r : per;
            const int k = next;
            next = (next + 1) % K;
            if ((rc = retire(k))) return rc;
            par_memcpy(c->pipe_pin[k], host + r * rowbytes, n * rowbytes);
            CU(cudaMemcpy2DAsync(dev + r * dpitch, dpitch, c->pipe_pin[k], rowbytes, rowbytes, n, cudaMemcpyHostToDevice, c->stream));
            CU(cudaEventRecord(c->pipe_ev[k], c->stream));
            slot[k].host = nullptr; slot[k].bytes = 0; slot[k].busy = true;
        }
        return 0;
    }
    int get(const char *dev, size_t dpitch, char *host, size_t rowbytes, size_t rows) {
        int rc = prepare(rowbytes);
        if (rc) return rc;
        const size_t per = c->pipe_bytes / rowbytes;
        for (size_t r = 0; r < rows; r += per) {
            const size_t n = rows - r < per ? rows - r : per;
            const int k = next;
            next = (next + 1) % K;
            if ((rc = retire(k))) return rc;
            CU(cudaMemcpy2DAsync(c->pipe_pin[k], rowbytes, dev + r * dpitch, dpitch, rowbytes, n, cudaMemcpyDeviceToHost, c->stream));
            CU(cudaEventRecord(c->pipe_ev[k], c->stream));
            slot[k].host = host + r * rowbytes; slot[k].bytes = n * rowbytes; slot[k].busy = true;
        }
        return 0;
    }
    int drain() {
        for (int i = 0; i < K; ++i) {
            int rc = retire((next + i) % K);
            if (rc) return rc;
        }
        return 0;
    }
};

// host rows [y0, y0+ny) of the local slab ([q][ny][nx]) -> device buffer b
static int upload_buffer(lbm_ctx *c, int b, const double *f, int y0 = 0, int ny = -1, bool sync = true) {
    const int nx = c->desc.nx, Q = c->li.Q;
    if (ny < 0) ny = c->nyl;
    if (ny == 0) return 0;
    if (is64(c)) {
        double *o = origin<double>(c, b) + (size_t)y0 * c->pitch;
        if (is_pinned(f)) {
            for (int i = 0; i < Q; ++i)
                CU(cudaMemcpy2DAsync(o + (size_t)i * c->plane, c->pitch * 8, f + (size_t)i * ny * nx, (size_t)nx * 8,
                                     (size_t)nx * 8, ny, cudaMemcpyHostToDevice, c->stream));
        } else {
            HostPipe pipe(c);
            for (int i = 0; i < Q; ++i) {
                int rc = pipe.put((char *)(o + (size_t)i * c->plane), c->pitch * 8, (const char *)(f + (size_t)i * ny * nx), (size_t)nx * 8, ny);
                if (rc) return rc;
            }
            int rc = pipe.drain();
            if (rc) return rc;
        }
    } else {
        int rc = need_stage(c, (size_t)ny * nx * 8);
        if (rc) return rc;
        double *stage = c->stage;
        KParams<float> p = make_params<float>(c, b, b);
        p.dst += (size_t)y0 * c->pitch;
        p.nyl = ny;
        const bool pinned = is_pinned(f);
        HostPipe pipe(c);
        for (int i = 0; i < Q; ++i) {
            if (pinned) {
                CU(cudaMemcpyAsync(stage, f + (size_t)i * ny * nx, (size_t)ny * nx * 8, cudaMemcpyHostToDevice, c->stream));
            } else {
                int rc2 = pipe.put((char *)stage, (size_t)nx * 8, (const char *)(f + (size_t)i * ny * nx), (size_t)nx * 8, ny);
                if (!rc2) rc2 = pipe.drain();
                if (rc2) return rc2;
            }
            c->ops->import32(p, stage, i, c->stream);
            c->launches += 1;
        }
    }
    if (sync) CU(cudaStreamSynchronize(c->stream));
    return 0;
}

int lbm_upload_f(lbm_ctx *c, const double *f) {
    if (!c || !f) return fail(LBM_ERR_INVALID, "null argument");
    CU(cudaSetDevice(c->desc.device));
    int rc = wait_comm(c);
    if (rc) return rc;
    rc = upload_buffer(c, c->cur, f);
    if (rc) return rc;
    c->state = ST_STREAM;
    c->have_coll = false;
    c->resume_ok = false;
    return 0;
}

int lbm_upload_f_collision(lbm_ctx *c, const double *f) {
    if (!c || !f) return fail(LBM_ERR_INVALID, "null argument");
    CU(cudaSetDevice(c->desc.device));
    int rc = materialize(c);
    if (rc) return rc;
    rc = wait_comm(c);
    if (rc) return rc;
    rc = upload_buffer(c, 1 - c->cur, f);
    if (rc) return rc;
    rc = refresh_ghosts(c, 1 - c->cur);
    if (rc) return rc;
    rc = post_exchange(c, 1 - c->cur);
    if (rc) return rc;
    c->have_coll = true;
    c->resume_ok = false;
    return 0;
}

static int download_buffer(lbm_ctx *c, int b, double *f, int y0 = 0, int ny = -1, bool sync = true) {
    const int nx = c->desc.nx, Q = c->li.Q;
    if (ny < 0) ny = c->nyl;
    if (ny == 0) return 0;
    if (is64(c)) {
        const double *o = origin<double>(c, b) + (size_t)y0 * c->pitch;
        if (is_pinned(f)) {
            for (int i = 0; i < Q; ++i)
                CU(cudaMemcpy2DAsync(f + (size_t)i * ny * nx, (size_t)nx * 8, o + (size_t)i * c->plane, c->pitch * 8,
                                     (size_t)nx * 8, ny, cudaMemcpyDeviceToHost, c->stream));
        } else {
            HostPipe pipe(c);
            for (int i = 0; i < Q; ++i) {
                int rc = pipe.get((const char *)(o + (size_t)i * c->plane), c->pitch * 8, (char *)(f + (size_t)i * ny * nx), (size_t)nx * 8, ny);
                if (rc) return rc;
            }
            int rc = pipe.drain();
            if (rc) return rc;
        }
    } else {
        int rc = need_stage(c, (size_t)ny * nx * 8);
        if (rc) return rc;
        double *stage = c->stage;
        KParams<float> p = make_params<float>(c, b, b);
        p.src += (size_t)y0 * c->pitch;
        p.nyl = ny;
        const bool pinned = is_pinned(f);
        HostPipe pipe(c);
        for (int i = 0; i < Q; ++i) {
            c->ops->export32(p, stage, i, c->stream);
            c->launches += 1;
            if (pinned) {
                CU(cudaMemcpyAsync(f + (size_t)i * ny * nx, stage, (size_t)ny * nx * 8, cudaMemcpyDeviceToHost, c->stream));
            } else {
                // (the single staging plane is rewritten by the next export: finish this plane's copy first)
                int rc2 = pipe.get((const char *)stage, (size_t)nx * 8, (char *)(f + (size_t)i * ny * nx), (size_t)nx * 8, ny);
                if (!rc2) rc2 = pipe.drain();
                if (rc2) return rc2;
            }
        }
    }
    if (sync) CU(cudaStreamSynchronize(c->stream));
    return 0;
}

int lbm_download_f(lbm_ctx *c, double *f) {
    if (!c || !f) return fail(LBM_ERR_INVALID, "null argument");
    CU(cudaSetDevice(c->desc.device));
    int rc = materialize(c);
    if (rc) return rc;
    rc = download_buffer(c, c->cur, f);
    return rc ? rc : p2p_check(c);
}

// Asynchronous forms for page-locked arrays: the copies are enqueued on the context's stream and the call returns; with two
// contexts in flight job k + 1's upload and job k - 1's download overlap job k's steps (PCIe is full duplex).  lbm_sync
// (or any synchronising call) completes them; `f` must stay valid and unmodified until then.
int lbm_upload_f_async(lbm_ctx *c, const double *f) {
    if (!c || !f) return fail(LBM_ERR_INVALID, "null argument");
    if (!is_pinned(f)) return fail(LBM_ERR_INVALID, "lbm_upload_f_async needs a page-locked array (lbm_host_alloc)");
    CU(cudaSetDevice(c->desc.device));
    int rc = wait_comm(c);
    if (rc) return rc;
    rc = upload_buffer(c, c->cur, f, 0, -1, false);
    if (rc) return rc;
    c->state = ST_STREAM;
    c->have_coll = false;
    c->resume_ok = false;
    return 0;
}

int lbm_download_f_async(lbm_ctx *c, double *f) {
    if (!c || !f) return fail(LBM_ERR_INVALID, "null argument");
    if (!is_pinned(f)) return fail(LBM_ERR_INVALID, "lbm_download_f_async needs a page-locked array (lbm_host_alloc)");
    CU(cudaSetDevice(c->desc.device));
    int rc = materialize(c);
    if (rc) return rc;
    return download_buffer(c, c->cur, f, 0, -1, false);
}

int lbm_host_alloc(void **ptr, size_t bytes) {
    if (!ptr) return fail(LBM_ERR_INVALID, "null argument");
    *ptr = nullptr;
    cudaError_t e = cudaHostAlloc(ptr, bytes ? bytes : 1, cudaHostAllocPortable);
    if (e != cudaSuccess) { cudaGetLastError(); return fail(LBM_ERR_NOMEM, "cudaHostAlloc(%zu bytes): %s", bytes, cudaGetErrorString(e)); }
    return 0;
}

int lbm_host_free(void *ptr) {
    if (!ptr) return 0;
    cudaError_t e = cudaFreeHost(ptr);
    if (e != cudaSuccess) { cudaGetLastError(); return fail(LBM_ERR_CUDA, "cudaFreeHost: %s", cudaGetErrorString(e)); }
    return 0;
}

static bool is_pinned(const void *p) {
    cudaPointerAttributes a;
    if (cudaPointerGetAttributes(&a, p) != cudaSuccess) { cudaGetLastError(); return false; }
    return a.type == cudaMemoryTypeHost;
}

int lbm_snapshot_end(lbm_ctx *c) {
    if (!c) return fail(LBM_ERR_INVALID, "null argument");
    if (!c->snap_pending) return 0;
    CU(cudaSetDevice(c->desc.device));
    c->snap_pending = false;
    CU(cudaEventSynchronize(c->ev_snap_done));
    if (!c->snap_direct) memcpy(c->snap_user, c->snap_host, (size_t)c->li.Q * c->nyl * c->desc.nx * 8);
    return p2p_check(c);
}

int lbm_snapshot_begin(lbm_ctx *c, double *f) {
    if (!c || !f) return fail(LBM_ERR_INVALID, "null argument");
    CU(cudaSetDevice(c->desc.device));
    int rc = lbm_snapshot_end(c);
    if (rc) return rc;
    rc = wait_comm(c);
    if (rc) return rc;
    const size_t bytes = (size_t)c->li.Q * c->nyl * c->desc.nx * 8;
    if (!c->copy_stream) {
        CU(cudaStreamCreateWithFlags(&c->copy_stream, cudaStreamNonBlocking));
        CU(cudaEventCreateWithFlags(&c->ev_snap, cudaEventDisableTiming));
        CU(cudaEventCreateWithFlags(&c->ev_snap_done, cudaEventDisableTiming));
    }
    if (c->snap_dev_bytes < bytes) {
        if (c->snap_dev) { cudaFree(c->snap_dev); c->snap_dev = nullptr; c->snap_dev_bytes = 0; }
        cudaError_t e = cudaMalloc(&c->snap_dev, bytes);
        if (e != cudaSuccess) { cudaGetLastError(); return fail(LBM_ERR_NOMEM, "cudaMalloc(%zu bytes for the snapshot buffer): %s", bytes, cudaGetErrorString(e)); }
        c->snap_dev_bytes = bytes;
    }
    c->snap_direct = is_pinned(f);
    if (!c->snap_direct && c->snap_host_bytes < bytes) {
        if (c->snap_host) { cudaFreeHost(c->snap_host); c->snap_host = nullptr; c->snap_host_bytes = 0; }
        cudaError_t e = cudaHostAlloc((void **)&c->snap_host, bytes, cudaHostAllocDefault);
        if (e != cudaSuccess) { cudaGetLastError(); return fail(LBM_ERR_NOMEM, "cudaHostAlloc(%zu bytes of snapshot staging): %s", bytes, cudaGetErrorString(e)); }
        c->snap_host_bytes = bytes;
    }
    const bool pull = c->state == ST_COLLIDED;
    if (is64(c)) c->ops->snapshot64(pull, make_params<double>(c, c->cur, c->cur), c->snap_dev, c->stream);
    else c->ops->snapshot32(pull, make_params<float>(c, c->cur, c->cur), c->snap_dev, c->stream);
    c->launches += 1;
    CU(cudaGetLastError());
    CU(cudaEventRecord(c->ev_snap, c->stream));
    CU(cudaStreamWaitEvent(c->copy_stream, c->ev_snap, 0));
    CU(cudaMemcpyAsync(c->snap_direct ? f : c->snap_host, c->snap_dev, bytes, cudaMemcpyDeviceToHost, c->copy_stream));
    CU(cudaEventRecord(c->ev_snap_done, c->copy_stream));
    c->snap_user = f;
    c->snap_pending = true;
    return 0;
}

int lbm_download_f_collision(lbm_ctx *c, double *f) {
    if (!c || !f) return fail(LBM_ERR_INVALID, "null argument");
    CU(cudaSetDevice(c->desc.device));
    if (c->state == ST_COLLIDED) {
        int rc = wait_comm(c);
        if (rc) return rc;
        rc = download_buffer(c, c->cur, f);
        return rc ? rc : p2p_check(c);
    }
    if (!c->have_coll) return fail(LBM_ERR_STATE, "no f_collision: collide! has not run since the last upload");
    int rc = download_buffer(c, 1 - c->cur, f);
    return rc ? rc : p2p_check(c);
}

static int check_rows(lbm_ctx *c, int y0, int ny) {
    if (y0 < 0 || ny < 0 || y0 + ny > c->nyl) return fail(LBM_ERR_INVALID, "rows [%d, %d) outside the local slab of %d rows", y0, y0 + ny, c->nyl);
    return 0;
}

int lbm_upload_f_rows(lbm_ctx *c, int32_t y0, int32_t ny, const double *f_rows) {
    if (!c || !f_rows) return fail(LBM_ERR_INVALID, "null argument");
    CU(cudaSetDevice(c->desc.device));
    int rc = check_rows(c, y0, ny);
    if (rc) return rc;
    rc = materialize(c);
    if (rc) return rc;
    rc = wait_comm(c);
    if (rc) return rc;
    rc = upload_buffer(c, c->cur, f_rows, y0, ny);
    if (rc) return rc;
    c->have_coll = false;
    c->resume_ok = false;
    return 0;
}

int lbm_init_equilibrium_rows(lbm_ctx *c, int32_t y0, int32_t ny, const double *rho, const double *ux, const double *uy,
                              const double *T) {
    if (!c || !rho || !ux || !uy || !T) return fail(LBM_ERR_INVALID, "null argument");
    CU(cudaSetDevice(c->desc.device));
    int rc = check_rows(c, y0, ny);
    if (rc) return rc;
    if (ny == 0) return 0;
    rc = materialize(c);
    if (rc) return rc;
    rc = wait_comm(c);
    if (rc) return rc;
    const size_t N = (size_t)ny * c->desc.nx;
    double *dev = nullptr;
    CU(cudaMalloc(&dev, 4 * N * 8));
    const double *host[4] = {rho, ux, uy, T};
    cudaError_t e = cudaSuccess;
    for (int k = 0; k < 4 && e == cudaSuccess; ++k)
        e = cudaMemcpyAsync(dev + k * N, host[k], N * 8, cudaMemcpyHostToDevice, c->stream);
    if (e == cudaSuccess) {
        if (is64(c)) {
            KParams<double> p = make_params<double>(c, c->cur, c->cur);
            p.dst += (size_t)y0 * c->pitch;
            p.nyl = ny;
            c->ops->init_eq64(p, dev, dev + N, dev + 2 * N, dev + 3 * N, c->stream);
        } else {
            KParams<float> p = make_params<float>(c, c->cur, c->cur);
            p.dst += (size_t)y0 * c->pitch;
            p.nyl = ny;
            c->ops->init_eq32(p, dev, dev + N, dev + 2 * N, dev + 3 * N, c->stream);
        }
        c->launches += 1;
        e = cudaGetLastError();
    }
    if (e == cudaSuccess) e = cudaStreamSynchronize(c->stream);
    cudaFree(dev);
    if (e != cudaSuccess) return fail(LBM_ERR_CUDA, "lbm_init_equilibrium_rows: %s", cudaGetErrorString(e));
    c->state = ST_STREAM;
    c->have_coll = false;
    c->resume_ok = false;
    return 0;
}

int lbm_init_analytic(lbm_ctx *c, const lbm_init_spec *spec) {
    if (!c || !spec) return fail(LBM_ERR_INVALID, "null argument");
    if (spec->offeq < 0 || spec->offeq > 2) return fail(LBM_ERR_INVALID, "offeq %d", spec->offeq);
    CU(cudaSetDevice(c->desc.device));
    int rc = materialize(c);
    if (rc) return rc;
    rc = wait_comm(c);
    if (rc) return rc;
    const int nx = c->desc.nx, nyl = c->nyl, W = nx + nyl;
    std::vector<double> tab((size_t)16 * W, 1.0);
    InitArgs ia;
    memset(&ia, 0, sizeof(ia));
    const lbm_sep_field *fields[8] = {&spec->rho, &spec->ux, &spec->uy, &spec->p, &spec->grad[0], &spec->grad[1], &spec->grad[2], &spec->grad[3]};
    for (int f = 0; f < 8; ++f) {
        ia.c0[f] = fields[f]->c0;
        for (int k = 0; k < 2; ++k) {
            ia.a[f][k] = fields[f]->a[k];
            double *t = tab.data() + (size_t)(2 * f + k) * W;
            if (fields[f]->x[k]) memcpy(t, fields[f]->x[k], (size_t)nx * 8);
            if (fields[f]->y[k]) memcpy(t + nx, fields[f]->y[k], (size_t)nyl * 8);
        }
    }
    ia.unit_density = spec->unit_density; ia.unit_temperature = spec->unit_temperature;
    ia.offeq = spec->offeq; ia.coef = spec->offeq_coef;
    double *dev = nullptr;
    CU(cudaMalloc(&dev, tab.size() * 8));
    cudaError_t e = cudaMemcpyAsync(dev, tab.data(), tab.size() * 8, cudaMemcpyHostToDevice, c->stream);
    ia.tab = dev;
    if (e == cudaSuccess) {
        if (is64(c)) c->ops->init_analytic64(make_params<double>(c, c->cur, c->cur), ia, c->stream);
        else c->ops->init_analytic32(make_params<float>(c, c->cur, c->cur), ia, c->stream);
        c->launches += 1;
        e = cudaGetLastError();
    }
    if (e == cudaSuccess) e = cudaStreamSynchronize(c->stream);
    cudaFree(dev);
    if (e != cudaSuccess) return fail(LBM_ERR_CUDA, "lbm_init_analytic: %s", cudaGetErrorString(e));
    c->state = ST_STREAM;
    c->have_coll = false;
    c->resume_ok = false;
    return 0;
}

int lbm_download_f_rows(lbm_ctx *c, int32_t y0, int32_t ny, double *f_rows) {
    if (!c || !f_rows) return fail(LBM_ERR_INVALID, "null argument");
    CU(cudaSetDevice(c->desc.device));
    int rc = check_rows(c, y0, ny);
    if (rc) return rc;
    rc = materialize(c);
    if (rc) return rc;
    rc = download_buffer(c, c->cur, f_rows, y0, ny);
    return rc ? rc : p2p_check(c);
}

static void free_force(lbm_ctx *c) {
    cudaStreamSynchronize(c->stream);
    drop_graphs(c);
    for (void **p : {&c->field, &c->sep_fx, &c->sep_fy})
        if (*p) { cudaFree(*p); *p = nullptr; }
    c->force_mode = 0;
}

// host double array -> device array of the context's element type
static int upload_as_elt(lbm_ctx *c, const double *h, size_t n, void **dev) {
    if (!*dev) CU(cudaMalloc(dev, n * c->elt));
    if (is64(c)) {
        CU(cudaMemcpy(*dev, h, n * 8, cudaMemcpyHostToDevice));
    } else {
        std::vector<float> tmp(n);
        for (size_t k = 0; k < n; ++k) tmp[k] = (float)h[k];
        CU(cudaMemcpy(*dev, tmp.data(), n * 4, cudaMemcpyHostToDevice));
    }
    return 0;
}

// per-node 2-component field (static force / prescribed velocity).  A host closure re-evaluated every step lands here once
// per step: the allocation (and the captured graphs, which only bake the pointer) are kept when the mode is unchanged.
static int set_node_field(lbm_ctx *c, const double *F) {
    if (c->force_mode == 2 && c->field) {
        CU(cudaStreamSynchronize(c->stream));
    } else {
        free_force(c);
    }
    int rc = upload_as_elt(c, F, (size_t)2 * c->nyl * c->desc.nx, &c->field);
    if (rc) return rc;
    c->force_mode = 2;
    return 0;
}

int lbm_set_force_none(lbm_ctx *c) {
    if (!c) return fail(LBM_ERR_INVALID, "null context");
    CU(cudaSetDevice(c->desc.device));
    free_force(c);
    return 0;
}

int lbm_set_force_uniform(lbm_ctx *c, double fx, double fy) {
    if (!c) return fail(LBM_ERR_INVALID, "null context");
    CU(cudaSetDevice(c->desc.device));
    if (c->desc.collision == LBM_MRT && c->desc.ntau < 2) return fail(LBM_ERR_INVALID, "MRT forcing needs tau[1] (mrt.jl:94)");
    if (c->desc.collision == LBM_ITERATIVE_INIT) return fail(LBM_ERR_UNSUPPORTED, "IterativeInitializationCollisionModel has no force (iterative_initialization.jl:1-12)");
    free_force(c);
    c->force_mode = 1; c->fx = fx; c->fy = fy;
    return 0;
}

int lbm_set_force_field(lbm_ctx *c, const double *F) {
    if (!c || !F) return fail(LBM_ERR_INVALID, "null argument");
    CU(cudaSetDevice(c->desc.device));
    if (c->desc.collision == LBM_MRT && c->desc.ntau < 2) return fail(LBM_ERR_INVALID, "MRT forcing needs tau[1] (mrt.jl:94)");
    if (c->desc.collision == LBM_ITERATIVE_INIT) return fail(LBM_ERR_UNSUPPORTED, "IterativeInitializationCollisionModel has no force (iterative_initialization.jl:1-12)");
    return set_node_field(c, F);
}

int lbm_set_velocity_field(lbm_ctx *c, const double *u0) {
    if (!c || !u0) return fail(LBM_ERR_INVALID, "null argument");
    if (c->desc.collision != LBM_ITERATIVE_INIT) return fail(LBM_ERR_STATE, "lbm_set_velocity_field is for LBM_ITERATIVE_INIT contexts");
    CU(cudaSetDevice(c->desc.device));
    return set_node_field(c, u0);
}

int lbm_set_force_separable(lbm_ctx *c, int64_t t0, int32_t nsteps, const double *fx_of_y, const double *fy_of_x) {
    if (!c || !fx_of_y || !fy_of_x || nsteps < 1) return fail(LBM_ERR_INVALID, "bad argument");
    CU(cudaSetDevice(c->desc.device));
    if (c->desc.collision == LBM_MRT && c->desc.ntau < 2) return fail(LBM_ERR_INVALID, "MRT forcing needs tau[1] (mrt.jl:94)");
    if (c->desc.collision == LBM_ITERATIVE_INIT) return fail(LBM_ERR_UNSUPPORTED, "IterativeInitializationCollisionModel has no force (iterative_initialization.jl:1-12)");
    free_force(c);
    int rc = upload_as_elt(c, fx_of_y, (size_t)nsteps * c->nyl, &c->sep_fx);
    if (rc) return rc;
    rc = upload_as_elt(c, fy_of_x, (size_t)nsteps * c->desc.nx, &c->sep_fy);
    if (rc) return rc;
    c->sep_t0 = t0; c->sep_n = nsteps;
    c->force_mode = 3;
    return 0;
}

static int check_force_window(lbm_ctx *c, int64_t t0, int64_t n) {
    if (c->desc.collision == LBM_ITERATIVE_INIT && c->force_mode != 2)
        return fail(LBM_ERR_STATE, "LBM_ITERATIVE_INIT needs the prescribed velocity: call lbm_set_velocity_field first");
    if (c->force_mode == 3 && (t0 < c->sep_t0 || t0 + n > c->sep_t0 + c->sep_n))
        return fail(LBM_ERR_STATE, "steps [%lld, %lld) outside the separable force table [%lld, %lld)", (long long)t0,
                    (long long)(t0 + n), c->sep_t0, c->sep_t0 + c->sep_n);
    return 0;
}

int lbm_collide(lbm_ctx *c, int64_t step, double time) {
    (void)time;  // time only enters through the force data
    if (!c) return fail(LBM_ERR_INVALID, "null context");
    CU(cudaSetDevice(c->desc.device));
    int rc = materialize(c);
    if (rc) return rc;
    rc = check_force_window(c, step, 1);
    if (rc) return rc;
    rc = is64(c) ? do_collide<double>(c, step) : do_collide<float>(c, step);
    if (rc) return rc;
    c->have_coll = true;
    c->resume_ok = false;
    return 0;
}

int lbm_stream(lbm_ctx *c) {
    if (!c) return fail(LBM_ERR_INVALID, "null context");
    CU(cudaSetDevice(c->desc.device));
    int rc = materialize(c);
    if (rc) return rc;
    if (!c->have_coll) return fail(LBM_ERR_STATE, "stream! needs f_collision: call lbm_collide first");
    rc = wait_comm(c);
    if (rc) return rc;
    if (is64(c)) {
        KParams<double> p = make_params<double>(c, 1 - c->cur, c->cur);
        p.nbc = 0; p.bc_sides = 0;
        c->ops->stream64(p, c->stream);
    } else {
        KParams<float> p = make_params<float>(c, 1 - c->cur, c->cur);
        p.nbc = 0; p.bc_sides = 0;
        c->ops->stream32(p, c->stream);
    }
    c->launches += 1;
    CU(cudaGetLastError());
    c->resume_ok = false;
    return 0;
}

int lbm_apply_bcs(lbm_ctx *c, double time) {
    (void)time;
    if (!c) return fail(LBM_ERR_INVALID, "null context");
    CU(cudaSetDevice(c->desc.device));
    int rc = materialize(c);
    if (rc) return rc;
    if (!c->have_coll) return fail(LBM_ERR_STATE, "apply! needs f_collision: call lbm_collide first");
    if (c->desc.n_bcs == 0) return 0;
    if (is64(c)) {
        KParams<double> p = make_params<double>(c, c->cur, c->cur);
        p.aux = origin<double>(c, 1 - c->cur);
        c->ops->bcs64(p, c->stream);
    } else {
        KParams<float> p = make_params<float>(c, c->cur, c->cur);
        p.aux = origin<float>(c, 1 - c->cur);
        c->ops->bcs32(p, c->stream);
    }
    c->launches += 1;
    CU(cudaGetLastError());
    return 0;
}

int lbm_step(lbm_ctx *c, int64_t t0, int64_t nsteps, double dt) {
    (void)dt;  // time = t*dt only enters through the force data (indexed by t)
    if (!c) return fail(LBM_ERR_INVALID, "null context");
    if (nsteps < 0) return fail(LBM_ERR_INVALID, "nsteps %lld", (long long)nsteps);
    if (nsteps == 0) return 0;
    CU(cudaSetDevice(c->desc.device));
    int rc = check_force_window(c, t0, nsteps);
    if (rc) return rc;
    if (c->p2p_failed || c->persist_failed) return p2p_check(c);
    CU(cudaEventRecord(c->ev_t0, c->stream));
    c->p2p_in_batch = false;
    if (c->p2p_on && c->opt_p2p) {
        // neighbour barrier: everything any of us enqueued before this batch (downloads, diagnostics, NCCL
        // exchanges) is finished before the first peer store of the batch
        ++c->batch;
        c->ops->p2p_barrier(my_flags(c), peer_flags(c->peer_up) + P2P_BATCH_FROM_DOWN, peer_flags(c->peer_dn) + P2P_BATCH_FROM_UP,
                            c->batch, c->stream);
        c->launches += 1;
    }
    rc = is64(c) ? do_steps<double>(c, t0, nsteps) : do_steps<float>(c, t0, nsteps);
    if (rc) return rc;
    if (c->desc.world > 1) {
        rc = wait_comm(c);  // the batch ends when the last halo has landed
        if (rc) return rc;
        if (c->p2p_in_batch) {
            c->ops->p2p_wait_epoch(my_flags(c), c->epoch, c->stream);
            c->launches += 1;
            CU(cudaGetLastError());
        }
    }
    CU(cudaEventRecord(c->ev_t1, c->stream));
    c->timed = true;
    return 0;
}

int lbm_sync(lbm_ctx *c) {
    if (!c) return fail(LBM_ERR_INVALID, "null context");
    CU(cudaSetDevice(c->desc.device));
    CU(cudaStreamSynchronize(c->stream));
    CU(cudaStreamSynchronize(c->bstream));
    CU(cudaStreamSynchronize(c->comm_stream));
    return p2p_check(c);
}

int lbm_last_step_ms(lbm_ctx *c, float *ms) {
    if (!c || !ms) return fail(LBM_ERR_INVALID, "null argument");
    if (!c->timed) return fail(LBM_ERR_STATE, "no lbm_step batch has run");
    CU(cudaSetDevice(c->desc.device));
    CU(cudaEventSynchronize(c->ev_t1));
    CU(cudaEventElapsedTime(ms, c->ev_t0, c->ev_t1));
    return 0;
}

int lbm_timer_start(lbm_ctx *c) {
    if (!c) return fail(LBM_ERR_INVALID, "null context");
    CU(cudaSetDevice(c->desc.device));
    CU(cudaEventRecord(c->ev_u0, c->stream));
    return 0;
}

int lbm_timer_stop(lbm_ctx *c, float *ms) {
    if (!c || !ms) return fail(LBM_ERR_INVALID, "null argument");
    CU(cudaSetDevice(c->desc.device));
    CU(cudaEventRecord(c->ev_u1, c->stream));
    CU(cudaEventSynchronize(c->ev_u1));
    CU(cudaEventElapsedTime(ms, c->ev_u0, c->ev_u1));
    return p2p_check(c);
}

int64_t lbm_kernel_launches(const lbm_ctx *c) { return c ? c->launches : 0; }
int lbm_halo_path(const lbm_ctx *c) { return !c || c->desc.world == 1 ? 0 : (c->p2p_on && c->opt_p2p ? 2 : 1); }

int lbm_set_option(lbm_ctx *c, const char *key, int64_t value) {
    if (!c || !key) return fail(LBM_ERR_INVALID, "null argument");
    CU(cudaSetDevice(c->desc.device));
    // "p2p" and "overlap" select the halo path: a collective choice (same value on every rank, before the next
    // lbm_step); a pending exchange of the old path is drained first
    int rc0 = wait_comm(c);
    if (rc0) return rc0;
    CU(cudaStreamSynchronize(c->stream));
    drop_graphs(c);
    if (!strcmp(key, "variant")) c->opt_variant = (int)value;
    else if (!strcmp(key, "graph")) c->opt_graph = (int)value;
    else if (!strcmp(key, "overlap")) c->opt_overlap = (int)value;
    else if (!strcmp(key, "p2p")) c->opt_p2p = (int)value;
    else if (!strcmp(key, "persistent")) c->opt_persistent = (int)value;
    else if (!strcmp(key, "prefetch")) c->opt_prefetch = (int)value;
    else if (!strcmp(key, "store")) c->opt_store = (int)value;
    else if (!strcmp(key, "tma")) c->opt_tma = (int)value;
    else if (!strcmp(key, "tma_cfg")) c->opt_tma_cfg = (int)value;
    else return fail(LBM_ERR_INVALID, "unknown option '%s'", key);
    return 0;
}

int lbm_moments(lbm_ctx *c, double tau_visc, double *rho, double *ux, double *uy, double *p, double *p_track,
                double *sxx, double *sxy, double *syy) {
    if (!c) return fail(LBM_ERR_INVALID, "null context");
    CU(cudaSetDevice(c->desc.device));
    int rc = wait_comm(c);
    if (rc) return rc;
    double *host[8] = {rho, ux, uy, p, p_track, sxx, sxy, syy};
    double *dev[8] = {nullptr};
    const size_t N = (size_t)c->nyl * c->desc.nx;
    int nf = 0;
    for (int k = 0; k < 8; ++k) nf += host[k] ? 1 : 0;
    // device fields live in a per-context buffer that is kept between calls (cudaMalloc / cudaFree of 100 MB-sized blocks
    // costs more than the kernel)
    if (c->mom_bytes < (size_t)nf * N * 8) {
        if (c->mom_dev) { CU(cudaStreamSynchronize(c->stream)); cudaFree(c->mom_dev); c->mom_dev = nullptr; c->mom_bytes = 0; }
        cudaError_t e = cudaMalloc(&c->mom_dev, (size_t)nf * N * 8);
        if (e != cudaSuccess) { cudaGetLastError(); return fail(LBM_ERR_NOMEM, "cudaMalloc(%zu): %s", (size_t)nf * N * 8, cudaGetErrorString(e)); }
        c->mom_bytes = (size_t)nf * N * 8;
    }
    for (int k = 0, j = 0; k < 8; ++k)
        if (host[k]) dev[k] = c->mom_dev + (size_t)(j++) * N;
    MomentsOut m{dev[0], dev[1], dev[2], dev[3], dev[4], dev[5], dev[6], dev[7], tau_visc};
    const bool pull = c->state == ST_COLLIDED;
    if (is64(c)) c->ops->moments64(pull, make_params<double>(c, c->cur, c->cur), m, c->stream);
    else c->ops->moments32(pull, make_params<float>(c, c->cur, c->cur), m, c->stream);
    c->launches += 1;
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return fail(LBM_ERR_CUDA, "lbm_moments: %s", cudaGetErrorString(e));
    {   // fields -> host: page-locked arrays directly, pageable ones through the chunk pipeline (several host threads)
        HostPipe pipe(c);
        const size_t rowbytes = (size_t)c->desc.nx * 8;
        for (int k = 0; k < 8; ++k) {
            if (!host[k]) continue;
            if (is_pinned(host[k])) { CU(cudaMemcpyAsync(host[k], dev[k], N * 8, cudaMemcpyDeviceToHost, c->stream)); }
            else { int rc2 = pipe.get((const char *)dev[k], rowbytes, (char *)host[k], rowbytes, (size_t)c->nyl); if (rc2) return rc2; }
        }
        int rc2 = pipe.drain();
        if (rc2) return rc2;
    }
    CU(cudaStreamSynchronize(c->stream));
    return p2p_check(c);
}

static int reduce_errors_mode(lbm_ctx *c, int mode, double tau_visc, double u_max, const lbm_sep_field *expected, double *out);
int lbm_reduce_errors(lbm_ctx *c, double tau_visc, double u_max, const lbm_sep_field *expected, double *out) {
    return reduce_errors_mode(c, 0, tau_visc, u_max, expected, out);
}
int lbm_reduce_process(lbm_ctx *c, double u_max, const lbm_sep_field *expected, double *out) {
    return reduce_errors_mode(c, 1, 1.0, u_max, expected, out);
}
// sum over the local nodes of E^2 for a separable field E(x, y) = c0 + sum_k a_k X_k(x) Y_k(y) (k < 2; a missing table
// is the constant 1): O(nx + ny) on the host instead of one more accumulator per node on the device
static double sep_square_sum(const lbm_sep_field &e, int nx, int ny) {
    double sx[2] = {0, 0}, sy[2] = {0, 0}, xx[2][2] = {{0, 0}, {0, 0}}, yy[2][2] = {{0, 0}, {0, 0}};
    auto X = [&](int k, int i) { return e.x[k] ? e.x[k][i] : 1.0; };
    auto Y = [&](int k, int j) { return e.y[k] ? e.y[k][j] : 1.0; };
    for (int k = 0; k < 2; ++k) {
        if (e.a[k] == 0.0) continue;
        for (int i = 0; i < nx; ++i) sx[k] += X(k, i);
        for (int j = 0; j < ny; ++j) sy[k] += Y(k, j);
        for (int l = k; l < 2; ++l) {
            if (e.a[l] == 0.0) continue;
            double a = 0, b = 0;
            for (int i = 0; i < nx; ++i) a += X(k, i) * X(l, i);
            for (int j = 0; j < ny; ++j) b += Y(k, j) * Y(l, j);
            xx[k][l] = xx[l][k] = a;
            yy[k][l] = yy[l][k] = b;
        }
    }
    double s = e.c0 * e.c0 * ((double)nx * (double)ny);
    for (int k = 0; k < 2; ++k) {
        if (e.a[k] == 0.0) continue;
        s += 2 * e.c0 * e.a[k] * sx[k] * sy[k];
        for (int l = 0; l < 2; ++l)
            if (e.a[l] != 0.0) s += e.a[k] * e.a[l] * xx[k][l] * yy[k][l];
    }
    return s;
}

static int reduce_errors_mode(lbm_ctx *c, int mode, double tau_visc, double u_max, const lbm_sep_field *expected, double *out) {
    if (!c || !expected || !out) return fail(LBM_ERR_INVALID, "null argument");
    if (!(tau_visc > 0) || !(u_max > 0)) return fail(LBM_ERR_INVALID, "tau_visc %g, u_max %g", tau_visc, u_max);
    CU(cudaSetDevice(c->desc.device));
    int rc = wait_comm(c);
    if (rc) return rc;
    const int nx = c->desc.nx, nyl = c->nyl, W = nx + nyl;
    const size_t tabn = (size_t)16 * W;
    const int nblocks = 8192;
    if (((long long)nx + 31) / 32 > nblocks) return fail(LBM_ERR_UNSUPPORTED, "reductions support at most %d columns", 32 * nblocks);
    const size_t need = tabn + (size_t)nblocks * 16 + 16;
    if (c->err_doubles < need) {
        if (c->err_dev) { CU(cudaStreamSynchronize(c->stream)); cudaFree(c->err_dev); c->err_dev = nullptr; c->err_doubles = 0; }
        CU(cudaMalloc(&c->err_dev, need * 8));
        c->err_doubles = need;
        c->err_tab.clear();
    }
    double *dev = c->err_dev;
    // Separable tables: a host copy of what the device holds is kept per (field, term) slot, and only slots whose
    // contents changed are uploaded -- between two calls of a run usually none (the time dependence of the analytic
    // fields sits in the coefficients), so the call is the two kernels plus a 128-byte read-back.  Slots of terms with a
    // zero coefficient are never read by the kernel and are skipped.
    if (c->err_tab.size() != tabn) { c->err_tab.assign(tabn, 0.0); c->err_tab_valid = 0; }
    ErrorArgs ea;
    memset(&ea, 0, sizeof(ea));
    cudaError_t e = cudaSuccess;
    for (int f = 0; f < 8; ++f) {
        ea.c0[f] = expected[f].c0;
        for (int k = 0; k < 2; ++k) {
            ea.a[f][k] = expected[f].a[k];
            if (expected[f].a[k] == 0.0) continue;
            const int slot = 2 * f + k;
            c->scratch_row.assign((size_t)W, 1.0);
            if (expected[f].x[k]) memcpy(c->scratch_row.data(), expected[f].x[k], (size_t)nx * 8);
            if (expected[f].y[k]) memcpy(c->scratch_row.data() + nx, expected[f].y[k], (size_t)nyl * 8);
            double *t = c->err_tab.data() + (size_t)slot * W;
            if (((c->err_tab_valid >> slot) & 1u) && memcmp(t, c->scratch_row.data(), (size_t)W * 8) == 0) continue;
            // (no launch is reading the device slot: every call of this function ends with a stream synchronise)
            memcpy(t, c->scratch_row.data(), (size_t)W * 8);
            if (e == cudaSuccess) e = cudaMemcpyAsync(dev + (size_t)slot * W, t, (size_t)W * 8, cudaMemcpyHostToDevice, c->stream);
            c->err_tab_valid |= 1u << slot;
        }
    }
    ea.tab = dev;
    ea.partials = dev + tabn;
    ea.out = ea.partials + (size_t)nblocks * 16;
    ea.nblocks = nblocks;
    ea.tau_visc = tau_visc;
    ea.u_max = u_max;
    ea.mode = mode;
    double h[16];
    if (e == cudaSuccess) {
        const bool pull = c->state == ST_COLLIDED;
        if (is64(c)) c->ops->errors64(pull, make_params<double>(c, c->cur, c->cur), ea, c->stream);
        else c->ops->errors32(pull, make_params<float>(c, c->cur, c->cur), ea, c->stream);
        c->launches += 2;
        e = cudaGetLastError();
    }
    if (e == cudaSuccess) e = cudaMemcpyAsync(h, ea.out, sizeof(h), cudaMemcpyDeviceToHost, c->stream);
    double sq[8] = {0};
    if (mode == 0)  // while the kernels run: the sums that involve the expected fields only
        for (int f = 1; f < 8; ++f) sq[f] = sep_square_sum(expected[f], nx, nyl);
    if (e == cudaSuccess) e = cudaStreamSynchronize(c->stream);
    if (e != cudaSuccess) return fail(LBM_ERR_CUDA, "lbm_reduce_errors: %s", cudaGetErrorString(e));
    if (mode == 0) { h[2] = sq[1] + sq[2]; h[4] = sq[3]; h[6] = sq[4]; h[8] = sq[5]; h[10] = sq[7]; h[12] = sq[6]; }
    memcpy(out, h, sizeof(h));
    return p2p_check(c);
}

int lbm_reduce(lbm_ctx *c, int32_t kind, double *out, int32_t n) {
    if (!c || !out || n < 1) return fail(LBM_ERR_INVALID, "bad argument");
    if (kind < LBM_REDUCE_MEAN_UX || kind > LBM_REDUCE_DENSITY_CHANGE) return fail(LBM_ERR_INVALID, "reduce kind %d", kind);
    CU(cudaSetDevice(c->desc.device));
    int rc = wait_comm(c);
    if (rc) return rc;
    // one partial per CTA, CTAs at most 256 columns wide: rows wider than 256 x red_blocks columns do not fit the buffer
    if (((long long)c->desc.nx + 31) / 32 > c->red_blocks) return fail(LBM_ERR_UNSUPPORTED, "reductions support at most %d columns", 32 * c->red_blocks);
    const size_t N = (size_t)c->nyl * c->desc.nx;
    if (kind == LBM_REDUCE_VELOCITY_CHANGE && !c->u_old) {
        CU(cudaMalloc(&c->u_old, 2 * N * 8));
        CU(cudaMemsetAsync(c->u_old, 0, 2 * N * 8, c->stream));  // zeros(T, 2) per node, stopping_criteria.jl:64
    }
    if (kind == LBM_REDUCE_DENSITY_CHANGE && !c->rho_old) {
        CU(cudaMalloc(&c->rho_old, N * 8));
        CU(cudaMemsetAsync(c->rho_old, 0, N * 8, c->stream));  // zeros(nx, ny), process_iterative_initialization.jl:9-10
    }
    ReduceArgs ra{kind, c->partials, kind == LBM_REDUCE_DENSITY_CHANGE ? c->rho_old : c->u_old, c->red_out, c->red_blocks, 0, 0};
    const bool pull = c->state == ST_COLLIDED;
    if (is64(c)) c->ops->reduce64(pull, make_params<double>(c, c->cur, c->cur), ra, c->stream);
    else c->ops->reduce32(pull, make_params<float>(c, c->cur, c->cur), ra, c->stream);
    c->launches += 2;
    CU(cudaGetLastError());
    double h[4];
    CU(cudaMemcpyAsync(h, c->red_out, sizeof(h), cudaMemcpyDeviceToHost, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    for (int k = 0; k < n && k < 4; ++k) out[k] = h[k];
    return p2p_check(c);
}


// ----------------------------------------------------------------------------------------------
// batched small problems (kernels: batch.cuh)
// ----------------------------------------------------------------------------------------------
}  // extern "C"

struct lbm_batch {
    lbm_desc desc;
    LatticeInfo li;
    const Ops *ops = nullptr;
    int nb = 0, N = 0;
    size_t elt = 8;
    void *f = nullptr;        // T [nb][Q][NY][NX]
    void *consts = nullptr;   // BatchConsts<T> [nb]
    std::vector<unsigned char> h_consts;  // host copy (relaxation times and force are set by separate calls)
    double *crit = nullptr;   // [nb][2 N]
    long long *steps = nullptr;
    int *stopped = nullptr;
    cudaStream_t stream = nullptr;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    bool timed = false;
    long long launches = 0;
};

template <typename T>
static BatchConsts<T> *consts_of(lbm_batch *b) { return reinterpret_cast<BatchConsts<T> *>(b->h_consts.data()); }

template <typename T>
static void batch_fill_tau(lbm_batch *b, const double *tau, int stride) {
    BatchConsts<T> *k = consts_of<T>(b);
    for (int i = 0; i < b->nb; ++i) fill_collision_consts<T>(k[i], b->desc.collision, tau + (size_t)i * stride, b->desc.ntau, b->li);
}

static int batch_upload_consts(lbm_batch *b) {
    CU(cudaMemcpyAsync(b->consts, b->h_consts.data(), b->h_consts.size(), cudaMemcpyHostToDevice, b->stream));
    CU(cudaStreamSynchronize(b->stream));
    return 0;
}

static int batch_check_range(const lbm_batch *b, int first, int count) {
    if (first < 0 || count < 0 || (long long)first + count > b->nb) return fail(LBM_ERR_INVALID, "problems [%d, %d) outside the batch of %d", first, first + count, b->nb);
    return 0;
}

__global__ void k_replicate(unsigned char *dst, const unsigned char *src, size_t bytes, long long copies) {
    // dst[c][i] = src[i], 16 bytes per thread where possible (bytes is a multiple of 4)
    const size_t total = bytes * (size_t)copies / 4;
    for (size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (size_t)gridDim.x * blockDim.x)
        reinterpret_cast<uint32_t *>(dst)[e] = reinterpret_cast<const uint32_t *>(src)[e % (bytes / 4)];
}

// reset the run state (step counter, stop flag, criterion memory) of problems [first, first + count)
static int batch_reset(lbm_batch *b, int first, int count) {
    CU(cudaMemsetAsync(b->steps + first, 0, (size_t)count * sizeof(long long), b->stream));
    CU(cudaMemsetAsync(b->stopped + first, 0, (size_t)count * sizeof(int), b->stream));
    CU(cudaMemsetAsync(b->crit + (size_t)first * 2 * b->N, 0, (size_t)count * 2 * b->N * sizeof(double), b->stream));
    return 0;
}

extern "C" {

void lbm_batch_destroy(lbm_batch *b) {
    if (!b) return;
    cudaSetDevice(b->desc.device);
    if (b->stream) cudaStreamSynchronize(b->stream);
    for (void *p : {b->f, b->consts, (void *)b->crit, (void *)b->steps, (void *)b->stopped})
        if (p) cudaFree(p);
    if (b->ev0) cudaEventDestroy(b->ev0);
    if (b->ev1) cudaEventDestroy(b->ev1);
    if (b->stream) cudaStreamDestroy(b->stream);
    delete b;
}

int lbm_batch_create(const lbm_desc *d, int32_t nbatch, lbm_batch **out) {
    if (!d || !out) return fail(LBM_ERR_INVALID, "null argument");
    *out = nullptr;
    LatticeInfo li;
    int rc = validate_desc(d, li);
    if (rc) return rc;
    if (nbatch < 1) return fail(LBM_ERR_INVALID, "nbatch %d", nbatch);
    if (d->world != 1) return fail(LBM_ERR_UNSUPPORTED, "a batch lives on one GPU: shard the problems over ranks instead (world must be 1)");
    if (d->collision == LBM_ITERATIVE_INIT) return fail(LBM_ERR_UNSUPPORTED, "LBM_ITERATIVE_INIT is not available in batches");
    const Ops *ops = get_ops(d->lattice, d->arith);
    if (!ops) return fail(LBM_ERR_UNSUPPORTED, "no kernels for lattice %d", d->lattice);
    const long long N = (long long)d->nx * d->ny;
    if (N * li.Q > 65535) return fail(LBM_ERR_UNSUPPORTED, "%d x %d nodes do not fit on chip: batches are for small problems, use one lbm_ctx per problem", d->nx, d->ny);
    lbm_batch *b = new lbm_batch();
    b->desc = *d; b->li = li; b->ops = ops; b->nb = nbatch; b->N = (int)N;
    b->elt = d->dtype == LBM_F64 ? 8 : 4;
    cudaError_t e = cudaSetDevice(d->device);
    if (e != cudaSuccess) { delete b; return fail(LBM_ERR_CUDA, "cudaSetDevice(%d): %s", d->device, cudaGetErrorString(e)); }
    auto bail = [&](int code) { lbm_batch_destroy(b); return code; };
    if (ops->init_constants() != 0) return bail(fail(LBM_ERR_CUDA, "constant upload failed: %s", cudaGetErrorString(cudaGetLastError())));
    const size_t csz = d->dtype == LBM_F64 ? sizeof(BatchConsts<double>) : sizeof(BatchConsts<float>);
    const size_t fbytes = (size_t)nbatch * li.Q * N * b->elt;
    if (cudaMalloc(&b->f, fbytes) != cudaSuccess || cudaMalloc(&b->consts, (size_t)nbatch * csz) != cudaSuccess ||
        cudaMalloc(&b->crit, (size_t)nbatch * 2 * N * sizeof(double)) != cudaSuccess ||
        cudaMalloc(&b->steps, (size_t)nbatch * sizeof(long long)) != cudaSuccess ||
        cudaMalloc(&b->stopped, (size_t)nbatch * sizeof(int)) != cudaSuccess)
        return bail(fail(LBM_ERR_NOMEM, "batch of %d problems (%zu bytes of populations): %s", nbatch, fbytes, cudaGetErrorString(cudaGetLastError())));
    if (cudaStreamCreateWithFlags(&b->stream, cudaStreamNonBlocking) != cudaSuccess || cudaEventCreate(&b->ev0) != cudaSuccess ||
        cudaEventCreate(&b->ev1) != cudaSuccess)
        return bail(fail(LBM_ERR_CUDA, "stream/event creation failed: %s", cudaGetErrorString(cudaGetLastError())));
    cudaMemsetAsync(b->f, 0, fbytes, b->stream);
    rc = batch_reset(b, 0, nbatch);
    if (rc) return bail(rc);
    // every problem starts with the descriptor's relaxation times and no force
    b->h_consts.assign((size_t)nbatch * csz, 0);
    std::vector<double> tau((size_t)nbatch * d->ntau);
    for (int i = 0; i < nbatch; ++i) memcpy(&tau[(size_t)i * d->ntau], d->tau, d->ntau * sizeof(double));
    if (d->dtype == LBM_F64) batch_fill_tau<double>(b, tau.data(), d->ntau);
    else batch_fill_tau<float>(b, tau.data(), d->ntau);
    rc = batch_upload_consts(b);
    if (rc) return bail(rc);
    *out = b;
    return 0;
}

int lbm_batch_set_tau(lbm_batch *b, const double *tau) {
    if (!b || !tau) return fail(LBM_ERR_INVALID, "null argument");
    CU(cudaSetDevice(b->desc.device));
    for (size_t i = 0; i < (size_t)b->nb * b->desc.ntau; ++i)
        if (!(tau[i] == tau[i]) || tau[i] == 0.0) return fail(LBM_ERR_INVALID, "tau[%zu][%zu] = %g", i / b->desc.ntau, i % b->desc.ntau, tau[i]);
    if (b->desc.dtype == LBM_F64) batch_fill_tau<double>(b, tau, b->desc.ntau);
    else batch_fill_tau<float>(b, tau, b->desc.ntau);
    return batch_upload_consts(b);
}

int lbm_batch_set_force_uniform(lbm_batch *b, const double *fxy) {
    if (!b) return fail(LBM_ERR_INVALID, "null argument");
    CU(cudaSetDevice(b->desc.device));
    if (fxy && b->desc.collision == LBM_MRT && b->desc.ntau < 2) return fail(LBM_ERR_INVALID, "MRT forcing needs tau[1] (mrt.jl:94)");
    for (int i = 0; i < b->nb; ++i) {
        if (b->desc.dtype == LBM_F64) {
            BatchConsts<double> &k = consts_of<double>(b)[i];
            k.forced = fxy != nullptr; k.fx = fxy ? fxy[2 * i] : 0; k.fy = fxy ? fxy[2 * i + 1] : 0;
        } else {
            BatchConsts<float> &k = consts_of<float>(b)[i];
            k.forced = fxy != nullptr; k.fx = fxy ? (float)fxy[2 * i] : 0; k.fy = fxy ? (float)fxy[2 * i + 1] : 0;
        }
    }
    return batch_upload_consts(b);
}

// host Float64 [count][q][ny][nx] <-> device storage (Float32 batches store f - w)
static int batch_copy_f(lbm_batch *b, int first, int count, double *f, bool upload) {
    const size_t per = (size_t)b->li.Q * b->N;
    if (count == 0) return 0;
    if (b->elt == 8) {
        double *dev = reinterpret_cast<double *>(b->f) + (size_t)first * per;
        CU(cudaMemcpyAsync(upload ? (void *)dev : (void *)f, upload ? (const void *)f : (const void *)dev, (size_t)count * per * 8,
                           upload ? cudaMemcpyHostToDevice : cudaMemcpyDeviceToHost, b->stream));
        CU(cudaStreamSynchronize(b->stream));
        return 0;
    }
    float *dev = reinterpret_cast<float *>(b->f) + (size_t)first * per;
    const size_t chunk = std::max<size_t>(1, ((size_t)8 << 20) / per);  // problems per staging chunk
    std::vector<float> tmp(std::min<size_t>(chunk, count) * per);
    for (size_t c0 = 0; c0 < (size_t)count; c0 += chunk) {
        const size_t n = std::min<size_t>(chunk, count - c0);
        if (upload) {
            for (size_t p = 0; p < n; ++p)
                for (int i = 0; i < b->li.Q; ++i)
                    for (int k = 0; k < b->N; ++k) tmp[(p * b->li.Q + i) * b->N + k] = (float)(f[((c0 + p) * b->li.Q + i) * b->N + k] - b->li.w[i]);
            CU(cudaMemcpyAsync(dev + c0 * per, tmp.data(), n * per * 4, cudaMemcpyHostToDevice, b->stream));
            CU(cudaStreamSynchronize(b->stream));
        } else {
            CU(cudaMemcpyAsync(tmp.data(), dev + c0 * per, n * per * 4, cudaMemcpyDeviceToHost, b->stream));
            CU(cudaStreamSynchronize(b->stream));
            for (size_t p = 0; p < n; ++p)
                for (int i = 0; i < b->li.Q; ++i)
                    for (int k = 0; k < b->N; ++k) f[((c0 + p) * b->li.Q + i) * b->N + k] = (double)tmp[(p * b->li.Q + i) * b->N + k] + b->li.w[i];
        }
    }
    return 0;
}

int lbm_batch_upload_f(lbm_batch *b, int32_t first, int32_t count, const double *f) {
    if (!b || !f) return fail(LBM_ERR_INVALID, "null argument");
    CU(cudaSetDevice(b->desc.device));
    int rc = batch_check_range(b, first, count);
    if (rc) return rc;
    rc = batch_copy_f(b, first, count, const_cast<double *>(f), true);
    return rc ? rc : batch_reset(b, first, count);
}

int lbm_batch_broadcast_f(lbm_batch *b, const double *f) {
    if (!b || !f) return fail(LBM_ERR_INVALID, "null argument");
    CU(cudaSetDevice(b->desc.device));
    int rc = batch_copy_f(b, 0, 1, const_cast<double *>(f), true);
    if (rc) return rc;
    if (b->nb > 1) {
        const size_t bytes = (size_t)b->li.Q * b->N * b->elt;
        k_replicate<<<2048, 256, 0, b->stream>>>((unsigned char *)b->f + bytes, (const unsigned char *)b->f, bytes, (long long)b->nb - 1);
        b->launches += 1;
        CU(cudaGetLastError());
    }
    return batch_reset(b, 0, b->nb);
}

int lbm_batch_download_f(lbm_batch *b, int32_t first, int32_t count, double *f) {
    if (!b || !f) return fail(LBM_ERR_INVALID, "null argument");
    CU(cudaSetDevice(b->desc.device));
    int rc = batch_check_range(b, first, count);
    if (rc) return rc;
    return batch_copy_f(b, first, count, f, false);
}

int lbm_batch_run(lbm_batch *b, int64_t nsteps, const lbm_batch_stop *stop) {
    if (!b) return fail(LBM_ERR_INVALID, "null argument");
    if (nsteps < 0) return fail(LBM_ERR_INVALID, "nsteps %lld", (long long)nsteps);
    CU(cudaSetDevice(b->desc.device));
    BatchParams p;
    memset(&p, 0, sizeof(p));
    p.nx = b->desc.nx; p.nyg = b->desc.ny; p.nb = b->nb;
    p.f = b->f; p.consts = b->consts; p.crit = b->crit; p.steps_done = b->steps; p.stopped = b->stopped;
    p.nsteps = nsteps;
    if (stop && stop->kind != LBM_BATCH_STOP_NONE) {
        if (stop->kind != LBM_BATCH_STOP_MEAN_UX && stop->kind != LBM_BATCH_STOP_VELOCITY_CHANGE) return fail(LBM_ERR_INVALID, "stop kind %d", stop->kind);
        if (stop->check_every < 1) return fail(LBM_ERR_INVALID, "check_every %d", stop->check_every);
        p.stop_kind = stop->kind; p.check_every = stop->check_every; p.tol = stop->tolerance;
    } else {
        p.check_every = 1;
    }
    p.nbc = b->desc.n_bcs;
    for (int k = 0; k < p.nbc; ++k) {
        const lbm_bc &s = b->desc.bcs[k];
        BCd &d = p.bc[k];
        d.kind = s.kind; d.dir = s.direction;
        d.x0 = s.x0; d.x1 = s.x1; d.y0 = s.y0; d.y1 = s.y1;
        d.ax = s.rho * s.u[0]; d.ay = s.rho * s.u[1];
        p.bc_sides |= 1 << s.direction;
        if (s.kind == LBM_BC_MOVING_WALL) p.has_mw = 1;
    }
    CU(cudaEventRecord(b->ev0, b->stream));
    const int rc = b->desc.dtype == LBM_F64 ? b->ops->batch64(b->desc.collision, p, b->stream) : b->ops->batch32(b->desc.collision, p, b->stream);
    if (rc == -1) return fail(LBM_ERR_UNSUPPORTED, "a %d x %d problem does not fit in shared memory", p.nx, p.nyg);
    if (rc) return fail(LBM_ERR_CUDA, "batch launch configuration failed: %s", cudaGetErrorString(cudaGetLastError()));
    CU(cudaGetLastError());
    CU(cudaEventRecord(b->ev1, b->stream));
    b->launches += 1;
    b->timed = true;
    return 0;
}

int lbm_batch_status(lbm_batch *b, int32_t first, int32_t count, int64_t *steps_done, int32_t *stopped) {
    if (!b) return fail(LBM_ERR_INVALID, "null argument");
    CU(cudaSetDevice(b->desc.device));
    int rc = batch_check_range(b, first, count);
    if (rc) return rc;
    static_assert(sizeof(long long) == sizeof(int64_t), "int64");
    if (steps_done) CU(cudaMemcpyAsync(steps_done, b->steps + first, (size_t)count * 8, cudaMemcpyDeviceToHost, b->stream));
    if (stopped) CU(cudaMemcpyAsync(stopped, b->stopped + first, (size_t)count * 4, cudaMemcpyDeviceToHost, b->stream));
    CU(cudaStreamSynchronize(b->stream));
    return 0;
}

int lbm_batch_reduce_errors(lbm_batch *b, const double *tau_visc, const double *u_max, const lbm_sep_field *expected,
                            const double *coef, double *out) {
    if (!b || !tau_visc || !u_max || !expected || !out) return fail(LBM_ERR_INVALID, "null argument");
    CU(cudaSetDevice(b->desc.device));
    const int nx = b->desc.nx, ny = b->desc.ny, W = nx + ny, nb = b->nb;
    for (int i = 0; i < nb; ++i)
        if (!(tau_visc[i] > 0) || !(u_max[i] > 0)) return fail(LBM_ERR_INVALID, "problem %d: tau_visc %g, u_max %g", i, tau_visc[i], u_max[i]);
    std::vector<double> tab((size_t)16 * W, 1.0), cf;
    for (int f = 0; f < 8; ++f)
        for (int k = 0; k < 2; ++k) {
            double *t = tab.data() + (size_t)(2 * f + k) * W;
            if (expected[f].x[k]) memcpy(t, expected[f].x[k], (size_t)nx * 8);
            if (expected[f].y[k]) memcpy(t + nx, expected[f].y[k], (size_t)ny * 8);
        }
    if (!coef) {  // the same coefficients for every problem
        cf.resize((size_t)nb * 24);
        for (int i = 0; i < nb; ++i)
            for (int f = 0; f < 8; ++f) { cf[(size_t)i * 24 + 3 * f] = expected[f].c0; cf[(size_t)i * 24 + 3 * f + 1] = expected[f].a[0]; cf[(size_t)i * 24 + 3 * f + 2] = expected[f].a[1]; }
        coef = cf.data();
    }
    double *dev = nullptr;
    const size_t n_in = tab.size() + (size_t)nb * (24 + 2), n_all = n_in + (size_t)nb * 16;
    CU(cudaMalloc(&dev, n_all * 8));
    double *d_tab = dev, *d_coef = d_tab + tab.size(), *d_tau = d_coef + (size_t)nb * 24, *d_um = d_tau + nb, *d_out = d_um + nb;
    cudaError_t e = cudaMemcpyAsync(d_tab, tab.data(), tab.size() * 8, cudaMemcpyHostToDevice, b->stream);
    if (e == cudaSuccess) e = cudaMemcpyAsync(d_coef, coef, (size_t)nb * 24 * 8, cudaMemcpyHostToDevice, b->stream);
    if (e == cudaSuccess) e = cudaMemcpyAsync(d_tau, tau_visc, (size_t)nb * 8, cudaMemcpyHostToDevice, b->stream);
    if (e == cudaSuccess) e = cudaMemcpyAsync(d_um, u_max, (size_t)nb * 8, cudaMemcpyHostToDevice, b->stream);
    if (e == cudaSuccess) {
        BatchErrorArgs ea{b->f, nx, ny, nb, d_tau, d_um, d_coef, d_tab, d_out};
        if (b->desc.dtype == LBM_F64) b->ops->batch_errors64(ea, b->stream);
        else b->ops->batch_errors32(ea, b->stream);
        b->launches += 1;
        e = cudaGetLastError();
    }
    if (e == cudaSuccess) e = cudaMemcpyAsync(out, d_out, (size_t)nb * 16 * 8, cudaMemcpyDeviceToHost, b->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(b->stream);
    cudaFree(dev);
    if (e != cudaSuccess) return fail(LBM_ERR_CUDA, "lbm_batch_reduce_errors: %s", cudaGetErrorString(e));
    return 0;
}

int lbm_batch_last_run_ms(lbm_batch *b, float *ms) {
    if (!b || !ms) return fail(LBM_ERR_INVALID, "null argument");
    if (!b->timed) return fail(LBM_ERR_STATE, "no lbm_batch_run has been issued");
    CU(cudaSetDevice(b->desc.device));
    CU(cudaEventSynchronize(b->ev1));
    CU(cudaEventElapsedTime(ms, b->ev0, b->ev1));
    return 0;
}

int64_t lbm_batch_kernel_launches(const lbm_batch *b) { return b ? b->launches : 0; }

}  // extern "C"
