// exaadmm_b200.cu — host side of the C ABI (include/exaadmm_b200.h): handle,
// HBM layout construction, launches, the fused inner loop and the native
// admm_two_level driver. No torch, no CPU fallback.
#include "../../include/exaadmm_b200.h"
#include "kernels.cuh"
#include "mp_kernels.cuh"
#include "qp_kernels.cuh"

#include <dlfcn.h>
#include <nccl.h>

#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <limits>
#include <string>
#include <mutex>
#include <vector>

using namespace ea;

namespace {

thread_local std::string g_create_error;

struct Timers { double x = 0, gen = 0, line = 0, bus = 0, z = 0, l = 0, lz = 0; };

}  // namespace

struct ea_handle {
    int device = 0;
    cudaStream_t stream = nullptr;
    int64_t ngen = 0, nline = 0, nbus = 0, nvar = 0;
    int nint = 0, gpad = 0;
    Dev d{};
    int zsel = 0;                               // host mirror of ctrl->zsel (valid whenever no fused run is in flight)
    std::vector<void *> allocs;                 // every cudaMalloc, freed in ea_destroy
    double *fields[EA_NUM_FIELDS] = { nullptr };    // device buffers; Z_CURR / Z_PREV resolved through zsel
    int *ref2int = nullptr;                     // nvar: HBM index of each reference-layout entry
    double *staging = nullptr;                  // max(nvar, nline, nbus) doubles (device)
    double *res_dev = nullptr;                  // 4 doubles (device) for step-wise norms
    Ctrl *ctrl_host = nullptr;                  // pinned
    double *res_host = nullptr;                 // pinned, 4 doubles
    std::vector<int> gen_of_slot;               // generator id (0-based, reference order) per slot
    std::vector<double> c2, c1, c0;             // reference order, for poststep
    int max_blocks = 0;
    branch::PowTable pow_table{};
    double pow_table_mu_max = -1.0;
    cudaEvent_t ev[4] = { nullptr, nullptr, nullptr, nullptr };
    Timers tm;
    int default_chunk = 16;
    // the chunk of the fused loop as a CUDA graph (every kernel argument is constant: what changes lives in ctrl)
    int use_graph = 1;
    cudaGraphExec_t graph = nullptr;
    int graph_chunk = 0, graph_max_auglag = 0, graph_count_work = 0;
    double graph_mu_max = 0.0, graph_scale = 0.0;
    long long n_replay = 0;
    int x_resident_blocks = 148;                // CTAs of k_xupdate resident on the device at once
    // launch accounting / optional per-kernel timing of the fused loop
    double span_s = 0.0;                        // device time spent inside ea_run_inner* calls (events on h->stream)
    long long n_x = 0, n_bus = 0, n_other = 0;  // kernels launched
    double t_x = 0.0, t_bus = 0.0;              // summed durations (kernel_timing only)
    long long n_timed = 0;                      // iterations those sums cover (no-op launches after `done` excluded)
    int kernel_timing = 0;
    void *flush_buf = nullptr;                  // option "l2_flush_mb": memset before every timed iteration (kernel_timing only)
    size_t flush_bytes = 0;
    int flush_clean = 0;                        // option "l2_flush_clean": then read a second buffer of the same size
    double *flush_sink = nullptr;
    int loopback = 0;                           // tests: exchange done by the caller through the host
    std::vector<cudaEvent_t> kev;               // event pool for kernel_timing
    cudaEvent_t span0 = nullptr, span1 = nullptr;
    // bus-partitioned multi-GPU mode
    int part_rank = 0, part_nranks = 1;
    int64_t n_owned_entries = 0;                // prefix of the HBM layout owned by this rank
    int64_t nvar_global = 0;                    // size of the whole problem (tolerances scale with sqrt of it)
    double *gather_dev = nullptr;               // nranks x stride
    double *gather_host = nullptr;              // pinned mirror (scalar collectives, loopback tests)
    ncclComm_t comm = nullptr;
    void *nccl_lib = nullptr;
    double *xbuf = nullptr;                     // cudaMalloc'ed exchange buffer (IPC-exported), peer mode
    std::vector<void *> peer_maps;              // cudaIpcOpenMemHandle'd peers
    std::string err;
};

namespace {
struct NcclApi {
    ncclResult_t (*GetUniqueId)(ncclUniqueId *) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*AllGather)(const void *, void *, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    const char *(*GetErrorString)(ncclResult_t) = nullptr;
    void *lib = nullptr;
};
NcclApi g_nccl;

// NCCL is loaded at run time (dlopen) so that the library has no link-time dependency on it:
// single-GPU users never touch it, and the process can share torch's bundled libnccl.
const char *load_nccl(const char *path) {
    if (g_nccl.lib) return nullptr;
    const char *cands[] = { path, getenv("EXAADMM_NCCL_LIB"), "libnccl.so.2", "libnccl.so" };
    void *lib = nullptr;
    for (const char *p : cands) {
        if (!p || !*p) continue;
        lib = dlopen(p, RTLD_NOW | RTLD_GLOBAL);
        if (lib) break;
    }
    if (!lib) return "cannot dlopen libnccl (pass its path or set EXAADMM_NCCL_LIB)";
    g_nccl.GetUniqueId = (decltype(g_nccl.GetUniqueId))dlsym(lib, "ncclGetUniqueId");
    g_nccl.CommInitRank = (decltype(g_nccl.CommInitRank))dlsym(lib, "ncclCommInitRank");
    g_nccl.AllGather = (decltype(g_nccl.AllGather))dlsym(lib, "ncclAllGather");
    g_nccl.CommDestroy = (decltype(g_nccl.CommDestroy))dlsym(lib, "ncclCommDestroy");
    g_nccl.GetErrorString = (decltype(g_nccl.GetErrorString))dlsym(lib, "ncclGetErrorString");
    if (!g_nccl.GetUniqueId || !g_nccl.CommInitRank || !g_nccl.AllGather || !g_nccl.CommDestroy || !g_nccl.GetErrorString)
        return "libnccl is missing a required symbol";
    g_nccl.lib = lib;
    return nullptr;
}
}  // namespace


namespace {

template <typename H> int fail(H *h, int code, const char *fmt, ...) {
    char buf[512];
    va_list ap; va_start(ap, fmt); vsnprintf(buf, sizeof(buf), fmt, ap); va_end(ap);
    if (h) h->err = buf; else g_create_error = buf;
    return code;
}

#define CK(call)                                                                                   \
    do {                                                                                           \
        cudaError_t e__ = (call);                                                                  \
        if (e__ != cudaSuccess)                                                                    \
            return fail(h, EA_ERR_CUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e__), __FILE__, __LINE__); \
    } while (0)

// Device memory comes from the stream-ordered pool (cudaMallocAsync) with an unlimited release threshold: handles
// are created and destroyed per solve by drop-in callers (solve_acopf builds a model per call), and the pool turns the
// ~50 allocations of ea_create into pointer bumps after the first handle.
template <typename H, typename T> int dev_alloc(H *h, T **p, size_t n) {
    void *q = nullptr;
    CK(cudaMallocAsync(&q, std::max<size_t>(n, 1) * sizeof(T), h->stream));
    CK(cudaMemsetAsync(q, 0, std::max<size_t>(n, 1) * sizeof(T), h->stream));
    h->allocs.push_back(q);
    *p = static_cast<T *>(q);
    return EA_OK;
}
template <typename H, typename T> int dev_upload(H *h, T **p, const std::vector<T> &v) {
    int rc = dev_alloc(h, p, v.size());
    if (rc) return rc;
    // v may be a temporary: a copy from pageable memory returns once the source has been read into the driver's staging
    // buffer (CUDA API synchronisation behaviour), so no stream synchronisation is needed before v goes away
    if (!v.empty()) CK(cudaMemcpyAsync(*p, v.data(), v.size() * sizeof(T), cudaMemcpyHostToDevice, h->stream));
    return EA_OK;
}

inline int nblocks(int64_t n, int block) { return (int)std::max<int64_t>(1, (n + block - 1) / block); }

// Upload straight from the caller's array (no host copy). A copy from pageable memory returns once the source has been
// read, so the caller's buffer may change afterwards.
template <typename H, typename T> int dev_upload_raw(H *h, T *dst, const T *src, size_t n) {
    if (n) CK(cudaMemcpyAsync(dst, src, n * sizeof(T), cudaMemcpyHostToDevice, h->stream));
    return EA_OK;
}

// Pinned blocks for the control-block mirror are recycled across handles: drop-in callers create a handle per solve, and
// a pinned allocation costs 1.3-2.5 ms each time.
static std::mutex g_pin_mutex;
static std::vector<void *> g_pin_free;
constexpr size_t PIN_BLOCK = ((sizeof(Ctrl) + 4 * sizeof(double) + 255) / 256) * 256;
static void *pin_acquire() {
    {
        std::lock_guard<std::mutex> lk(g_pin_mutex);
        if (!g_pin_free.empty()) { void *p = g_pin_free.back(); g_pin_free.pop_back(); return p; }
    }
    void *p = nullptr;
    return cudaMallocHost(&p, PIN_BLOCK) == cudaSuccess ? p : nullptr;
}
static void pin_release(void *p) {
    std::lock_guard<std::mutex> lk(g_pin_mutex);
    if (g_pin_free.size() < 64) g_pin_free.push_back(p); else cudaFreeHost(p);
}

void build_pow_table(ea_handle *h, double mu_max) {
    if (h->pow_table_mu_max == mu_max) return;
    branch::PowTable &T = h->pow_table;
    T.n = 0;
    double mu = 10.0;                         // acopf_auglag_linelimit_kernel_gpu.jl:75-77
    for (int k = 0; k < 24; ++k) {
        T.mu[k] = mu; T.inv_p01[k] = 1.0 / std::pow(mu, 0.1); T.p09[k] = std::pow(mu, 0.9);
        T.n = k + 1;
        const double next = std::min(mu_max, mu * 10);
        if (next == mu) break;
        mu = next;
    }
    h->pow_table_mu_max = mu_max;
}

double *field_ptr(ea_handle *h, int field) {
    if (field == EA_Z_CURR) return h->d.zbuf[h->zsel];
    if (field == EA_Z_PREV) return h->d.zbuf[h->zsel ^ 1];
    return h->fields[field];
}

int sync_ctrl_to_device(ea_handle *h, double beta, double eps_pri, long long inner0, long long limit) {
    k_ctrl_begin<<<1, 1, 0, h->stream>>>(h->d.ctrl, beta, eps_pri, inner0, limit, h->zsel);
    CK(cudaGetLastError());
    h->n_other++;
    return EA_OK;
}

int launch_x(ea_handle *h, long long major, int zsel, int max_auglag, double mu_max, double scale, int lines, int gens,
             cudaStream_t stream = nullptr) {
    if (!stream) stream = h->stream;
    build_pow_table(h, mu_max);
    if (major > 0 && lines)     // step-wise call: the fused loop resets the work queue itself (k_bus / k_ctrl_begin)
        CK(cudaMemsetAsync(&h->d.ctrl->next_line, 0, sizeof(int), stream));
    // persistent grid: as many CTAs as are resident at once, capped by the work available
    // one lane per branch at least: small grids are spread over all resident warps (k_xupdate: lanes_on)
    const int64_t work_blocks = std::max<int64_t>((h->nline + XBLOCK / 32 - 1) / (XBLOCK / 32), (h->ngen + XBLOCK - 1) / XBLOCK);
    const int grid = (int)std::max<int64_t>(1, std::min<int64_t>(h->x_resident_blocks, work_blocks));
    if (h->d.count_work)
        k_xupdate<true><<<grid, XBLOCK, XTILE_BYTES, stream>>>(h->d, h->pow_table, major, zsel, max_auglag, mu_max, scale, lines, gens);
    else
        k_xupdate<false><<<grid, XBLOCK, XTILE_BYTES, stream>>>(h->d, h->pow_table, major, zsel, max_auglag, mu_max, scale, lines, gens);
    CK(cudaGetLastError());
    h->n_x++;
    return EA_OK;
}

// Pack consecutive buses into warps of the bus kernel: the ends of a warp's buses fill at most 32 lanes; a bus
// with more than 32 ends (or none) gets a warp of its own and takes the scalar path. The result is one int4 per
// (warp, lane): {end slot or -1, bus, lane - leader lane, #ends of the bus if this lane leads it}.
int build_bus_warps(ea_handle *h, const std::vector<int> &hstart, int nbus_active) {
    std::vector<int4> info;
    auto flush = [&](std::vector<int4> &w) { w.resize(32, make_int4(-1, 0, 0, 0)); info.insert(info.end(), w.begin(), w.end()); w.clear(); };
    std::vector<int4> cur;
    for (int b = 0; b < nbus_active; ++b) {
        const int n = hstart[b + 1] - hstart[b];
        if (n > 32 || n == 0) {
            if (!cur.empty()) flush(cur);
            cur.push_back(make_int4(-1, b, 0, -1));
            flush(cur);
            continue;
        }
        if ((int)cur.size() + n > 32) flush(cur);
        for (int j = 0; j < n; ++j) cur.push_back(make_int4(hstart[b] + j, b, j, j == 0 ? n : 0));
    }
    if (!cur.empty()) flush(cur);
    h->d.n_bus_warps = (int)(info.size() / 32);
    return dev_upload(h, const_cast<int4 **>(&h->d.lane_info), info);
}

void drop_loop_graph(ea_handle *h) {
    if (h->graph) { cudaGraphExecDestroy(h->graph); h->graph = nullptr; }
}

double elapsed_s(cudaEvent_t a, cudaEvent_t b) { float ms = 0; cudaEventElapsedTime(&ms, a, b); return 1e-3 * ms; }

}  // namespace

// ===========================================================================
extern "C" {

int ea_abi_version(void) { return EA_ABI_VERSION; }

const char *ea_last_error(const ea_handle_t *h) { return h ? h->err.c_str() : g_create_error.c_str(); }

int ea_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
    return n;
}

int ea_create(const ea_grid_t *G, int device, ea_handle_t **out) {
    ea_handle *h = nullptr;
    if (!G || !out) return fail(h, EA_ERR_ARG, "ea_create: NULL argument");
    if (G->ngen < 0 || G->nline < 0 || G->nbus <= 0) return fail(h, EA_ERR_ARG, "ea_create: bad sizes");
    if (2 * G->ngen + 8 * G->nline > (int64_t)std::numeric_limits<int>::max() / 2)
        return fail(h, EA_ERR_ARG, "ea_create: problem too large for 32-bit indexing");
    int ndev = ea_device_count();
    if (ndev <= 0) return fail(h, EA_ERR_CUDA, "ea_create: no CUDA device available (this path has no CPU fallback)");
    if (device < 0 || device >= ndev) return fail(h, EA_ERR_ARG, "ea_create: device %d out of range (0..%d)", device, ndev - 1);

    h = new ea_handle();
    auto bail = [&](int rc) { g_create_error = h->err; ea_destroy(h); return rc; };
    const bool prof_create = getenv("EXAADMM_PROFILE_CREATE") != nullptr;
    auto lap_t = std::chrono::steady_clock::now();
    auto lap = [&](const char *what) {
        if (!prof_create) return;
        const auto now = std::chrono::steady_clock::now();
        fprintf(stderr, "[ea_create] %-28s %8.3f ms\n", what, 1e3 * std::chrono::duration<double>(now - lap_t).count());
        lap_t = now;
    };
    h->device = device;
    if (cudaSetDevice(device) != cudaSuccess) return bail(fail(h, EA_ERR_CUDA, "cudaSetDevice(%d) failed", device));
    if (cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking) != cudaSuccess)
        return bail(fail(h, EA_ERR_CUDA, "cudaStreamCreate failed"));
    {
        cudaMemPool_t pool;
        unsigned long long keep = ~0ull;
        if (cudaDeviceGetDefaultMemPool(&pool, device) == cudaSuccess)
            cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
    }
    for (auto &e : h->ev) if (cudaEventCreate(&e) != cudaSuccess) return bail(fail(h, EA_ERR_CUDA, "cudaEventCreate failed"));
    if (cudaEventCreate(&h->span0) != cudaSuccess || cudaEventCreate(&h->span1) != cudaSuccess)
        return bail(fail(h, EA_ERR_CUDA, "cudaEventCreate failed"));

    {
        int per_sm = 0, sms = 0;
        if (cudaFuncSetAttribute(k_xupdate<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, XTILE_BYTES) != cudaSuccess ||
            cudaFuncSetAttribute(k_xupdate<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, XTILE_BYTES) != cudaSuccess ||
            cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_xupdate<true>, XBLOCK, XTILE_BYTES) != cudaSuccess ||
            cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device) != cudaSuccess || per_sm < 1)
            return bail(fail(h, EA_ERR_CUDA, "k_xupdate occupancy query failed: %s", cudaGetErrorString(cudaGetLastError())));
        h->x_resident_blocks = per_sm * sms;
    }
    lap("stream, events, occupancy");
    const int ngen = (int)G->ngen, nline = (int)G->nline, nbus = (int)G->nbus;
    h->ngen = ngen; h->nline = nline; h->nbus = nbus; h->nvar = 2 * (int64_t)ngen + 8 * (int64_t)nline;
    const int gpad = ((2 * ngen + 3) / 4) * 4;
    const int nint = gpad + 8 * nline;
    h->gpad = gpad; h->nint = nint;

    // ---- validate and build the bus-sorted layout ----------------------------------
    std::vector<int> hstart(nbus + 1), gstart(nbus + 1), slot_from(nline, -1), slot_to(nline, -1), slot_of_gen(ngen, -1);
    h->gen_of_slot.assign(ngen, -1);
    for (int b = 0; b <= nbus; ++b) {
        const int64_t fs = G->FrStart[b] - 1, ts = G->ToStart[b] - 1, gs = G->GenStart[b] - 1;
        if (fs < 0 || fs > nline || ts < 0 || ts > nline || gs < 0 || gs > ngen)
            return bail(fail(h, EA_ERR_ARG, "ea_create: CSR pointer out of range at bus %d", b));
        if (b > 0 && (G->FrStart[b] < G->FrStart[b - 1] || G->ToStart[b] < G->ToStart[b - 1] || G->GenStart[b] < G->GenStart[b - 1]))
            return bail(fail(h, EA_ERR_ARG, "ea_create: CSR pointers are not monotonic at bus %d", b));
        hstart[b] = (int)(fs + ts);
        gstart[b] = (int)gs;
    }
    if (G->FrStart[0] != 1 || G->ToStart[0] != 1 || G->GenStart[0] != 1)
        return bail(fail(h, EA_ERR_ARG, "ea_create: CSR pointers must start at 1"));
    if (hstart[nbus] != 2 * nline || gstart[nbus] != ngen)
        return bail(fail(h, EA_ERR_ARG, "ea_create: CSR pointers do not cover all lines/generators"));
    for (int b = 0; b < nbus; ++b) {
        int s = hstart[b];
        for (int64_t k = G->FrStart[b] - 1; k < G->FrStart[b + 1] - 1; ++k, ++s) {
            const int64_t l = G->FrIdx[k] - 1;
            if (l < 0 || l >= nline || slot_from[l] >= 0) return bail(fail(h, EA_ERR_ARG, "ea_create: bad FrIdx entry %lld", (long long)k));
            slot_from[l] = s;
        }
        for (int64_t k = G->ToStart[b] - 1; k < G->ToStart[b + 1] - 1; ++k, ++s) {
            const int64_t l = G->ToIdx[k] - 1;
            if (l < 0 || l >= nline || slot_to[l] >= 0) return bail(fail(h, EA_ERR_ARG, "ea_create: bad ToIdx entry %lld", (long long)k));
            slot_to[l] = s;
        }
        for (int64_t k = G->GenStart[b] - 1; k < G->GenStart[b + 1] - 1; ++k) {
            const int64_t g = G->GenIdx[k] - 1;
            if (g < 0 || g >= ngen || slot_of_gen[g] >= 0) return bail(fail(h, EA_ERR_ARG, "ea_create: bad GenIdx entry %lld", (long long)k));
            slot_of_gen[g] = (int)k;
            h->gen_of_slot[k] = (int)g;
        }
    }
    for (int l = 0; l < nline; ++l)
        if (slot_from[l] < 0 || slot_to[l] < 0) return bail(fail(h, EA_ERR_ARG, "ea_create: line %d is missing from FrIdx / ToIdx", l));
    for (int g = 0; g < ngen; ++g)
        if (slot_of_gen[g] < 0) return bail(fail(h, EA_ERR_ARG, "ea_create: generator %d is missing from GenIdx", g));
    // (the index map reference layout -> HBM layout, nvar entries, is derived from these on the device: k_build_ref2int)

    lap("layout (host)");
    // ---- device allocations -------------------------------------------------------------
    int rc;
    Dev &d = h->d;
    d.ngen = ngen; d.nline = nline; d.nbus = nbus; d.nint = nint; d.gpad = gpad; d.baseMVA = G->baseMVA;
    for (int f = 0; f < EA_NUM_FIELDS; ++f)
        if ((rc = dev_alloc(h, &h->fields[f], (size_t)nint))) return bail(rc);
    d.u = h->fields[EA_U_CURR]; d.v = h->fields[EA_V_CURR]; d.l = h->fields[EA_L_CURR]; d.rho = h->fields[EA_RHO];
    d.lz = h->fields[EA_LZ]; d.zbuf[0] = h->fields[EA_Z_CURR]; d.zbuf[1] = h->fields[EA_Z_PREV];
    d.rp = h->fields[EA_RP]; d.rd = h->fields[EA_RD]; d.axby = h->fields[EA_AX_PLUS_BY];
    h->zsel = 0;

    lap("state vectors (alloc + memset)");
    if ((rc = dev_upload(h, const_cast<int **>(&d.slot_from), slot_from))) return bail(rc);
    if ((rc = dev_upload(h, const_cast<int **>(&d.slot_to), slot_to))) return bail(rc);
    if ((rc = dev_upload(h, const_cast<int **>(&d.hstart), hstart))) return bail(rc);
    if ((rc = dev_upload(h, const_cast<int **>(&d.gstart), gstart))) return bail(rc);
    {
        int *sog = nullptr;
        if ((rc = dev_upload(h, &sog, slot_of_gen))) return bail(rc);
        if ((rc = dev_alloc(h, &h->ref2int, (size_t)h->nvar))) return bail(rc);
        k_build_ref2int<<<nblocks(std::max(ngen, nline), 256), 256, 0, h->stream>>>(ngen, nline, gpad, sog, d.slot_from, d.slot_to, h->ref2int);
        if (cudaGetLastError() != cudaSuccess) return bail(fail(h, EA_ERR_CUDA, "k_build_ref2int launch failed"));
    }
    if ((rc = build_bus_warps(h, hstart, nbus))) return bail(rc);

    lap("index maps + bus warps");
    {   // per line: the caller's arrays go up as they are; the bounds (interleaved lower / upper pairs) and the bus
        // indices (1-based int64 pairs) are re-laid out on the device
        for (int l = 0; l < nline; ++l) {
            const int64_t fb = G->brBusIdx[2 * l] - 1, tb = G->brBusIdx[2 * l + 1] - 1;
            if (fb < 0 || fb >= nbus || tb < 0 || tb >= nbus) return bail(fail(h, EA_ERR_ARG, "ea_create: bad brBusIdx at line %d", l));
        }
        double *Yd = nullptr, *xlud = nullptr, *rated = nullptr, *raw = nullptr;
        int *brfd = nullptr, *brtd = nullptr;
        long long *idx = nullptr;
        if ((rc = dev_alloc(h, &Yd, 8 * (size_t)nline)) || (rc = dev_alloc(h, &xlud, 8 * (size_t)nline)) ||
            (rc = dev_alloc(h, &rated, (size_t)nline)) || (rc = dev_alloc(h, &brfd, (size_t)nline)) ||
            (rc = dev_alloc(h, &brtd, (size_t)nline)) || (rc = dev_alloc(h, &raw, 8 * (size_t)nline)) ||
            (rc = dev_alloc(h, &idx, 2 * (size_t)nline))) return bail(rc);
        const double *ys[8] = { G->YffR, G->YffI, G->YftR, G->YftI, G->YttR, G->YttI, G->YtfR, G->YtfI };
        for (int k = 0; k < 8; ++k)
            if ((rc = dev_upload_raw(h, Yd + (size_t)k * nline, ys[k], (size_t)nline))) return bail(rc);
        const double *bs[4] = { G->FrVmBound, G->ToVmBound, G->FrVaBound, G->ToVaBound };
        for (int k = 0; k < 4; ++k)
            if ((rc = dev_upload_raw(h, raw + (size_t)k * 2 * nline, bs[k], 2 * (size_t)nline))) return bail(rc);
        if ((rc = dev_upload_raw(h, rated, G->rateA, (size_t)nline))) return bail(rc);
        static_assert(sizeof(long long) == sizeof(int64_t), "brBusIdx is uploaded as 64-bit integers");
        if ((rc = dev_upload_raw(h, idx, reinterpret_cast<const long long *>(G->brBusIdx), 2 * (size_t)nline))) return bail(rc);
        k_layout_lines<<<nblocks(nline, 256), 256, 0, h->stream>>>(nline, raw, idx, xlud, brfd, brtd);
        if (cudaGetLastError() != cudaSuccess) return bail(fail(h, EA_ERR_CUDA, "k_layout_lines launch failed"));
        d.Y = Yd; d.xlu = xlud; d.rateA = rated; d.br_from = brfd; d.br_to = brtd;
        if ((rc = dev_alloc(h, &d.als, 3 * (size_t)nline))) return bail(rc);       // membuf rows 25-27 start at 0 (acopf_model.jl:87-88)
    }
    lap("per-line data");
    {   // per generator slot
        auto permute = [&](const double *src) {
            std::vector<double> v(ngen);
            for (int k = 0; k < ngen; ++k) v[k] = src[h->gen_of_slot[k]];
            return v;
        };
        if ((rc = dev_upload(h, const_cast<double **>(&d.pgmin), permute(G->pgmin)))) return bail(rc);
        if ((rc = dev_upload(h, const_cast<double **>(&d.pgmax), permute(G->pgmax)))) return bail(rc);
        if ((rc = dev_upload(h, const_cast<double **>(&d.pgmin_curr), permute(G->pgmin)))) return bail(rc);   // acopf_model.jl:61-64
        if ((rc = dev_upload(h, const_cast<double **>(&d.pgmax_curr), permute(G->pgmax)))) return bail(rc);
        if ((rc = dev_upload(h, const_cast<double **>(&d.qgmin), permute(G->qgmin)))) return bail(rc);
        if ((rc = dev_upload(h, const_cast<double **>(&d.qgmax), permute(G->qgmax)))) return bail(rc);
        if ((rc = dev_upload(h, const_cast<double **>(&d.c2), permute(G->c2)))) return bail(rc);
        if ((rc = dev_upload(h, const_cast<double **>(&d.c1), permute(G->c1)))) return bail(rc);
        h->c2.assign(G->c2, G->c2 + ngen); h->c1.assign(G->c1, G->c1 + ngen); h->c0.assign(G->c0, G->c0 + ngen);
    }
    {   // per bus
        std::vector<double> pd(nbus), qd(nbus);
        for (int b = 0; b < nbus; ++b) { pd[b] = G->Pd[b] / G->baseMVA; qd[b] = G->Qd[b] / G->baseMVA; }   // acopf_bus_kernel_gpu.jl:64-65
        if ((rc = dev_upload(h, const_cast<double **>(&d.pd_pu), pd))) return bail(rc);
        if ((rc = dev_upload(h, const_cast<double **>(&d.qd_pu), qd))) return bail(rc);
        if ((rc = dev_upload(h, const_cast<double **>(&d.YshR), std::vector<double>(G->YshR, G->YshR + nbus)))) return bail(rc);
        if ((rc = dev_upload(h, const_cast<double **>(&d.YshI), std::vector<double>(G->YshI, G->YshI + nbus)))) return bail(rc);
        if ((rc = dev_upload(h, const_cast<double **>(&d.Vmin), std::vector<double>(G->Vmin, G->Vmin + nbus)))) return bail(rc);
        if ((rc = dev_upload(h, const_cast<double **>(&d.Vmax), std::vector<double>(G->Vmax, G->Vmax + nbus)))) return bail(rc);
    }
    lap("per-generator / per-bus data");
    h->max_blocks = std::max(nblocks((int64_t)2 * nline + 32 * (int64_t)nbus, BBLOCK), 1024);
    if ((rc = dev_alloc(h, &d.partials, 4 * (size_t)h->max_blocks))) return bail(rc);
    if ((rc = dev_alloc(h, &d.ctrl, 1))) return bail(rc);
    if ((rc = dev_alloc(h, &d.counters, 1))) return bail(rc);
    if ((rc = dev_alloc(h, &h->res_dev, 4))) return bail(rc);
    if ((rc = dev_alloc(h, &h->staging, (size_t)std::max<int64_t>({ h->nvar, (int64_t)nline, (int64_t)nbus, (int64_t)ngen, 1 })))) return bail(rc);
    d.count_work = 1;
    d.nbus_active = nbus;
    h->n_owned_entries = nint;
    h->nvar_global = h->nvar;
    {   // one pinned block for the control-block mirror and the 4 norms (a pinned allocation costs ~1.3 ms)
        static_assert(sizeof(Ctrl) % sizeof(double) == 0, "res_host follows ctrl_host in one pinned block");
        void *pin = pin_acquire();
        if (!pin) return bail(fail(h, EA_ERR_ALLOC, "cudaMallocHost failed"));
        h->ctrl_host = static_cast<Ctrl *>(pin);
        h->res_host = reinterpret_cast<double *>(static_cast<char *>(pin) + sizeof(Ctrl));
    }
    memset(h->ctrl_host, 0, sizeof(Ctrl));
    lap("control blocks, pinned memory");
    *out = h;
    return EA_OK;
}

void ea_destroy(ea_handle_t *h) {
    if (!h) return;
    cudaSetDevice(h->device);
    if (h->stream) cudaStreamSynchronize(h->stream);
    drop_loop_graph(h);
    if (h->comm && g_nccl.CommDestroy) g_nccl.CommDestroy(h->comm);
    for (void *p : h->peer_maps) if (p) cudaIpcCloseMemHandle(p);
    if (h->xbuf) cudaFree(h->xbuf);
    if (h->flush_buf) cudaFree(h->flush_buf);
    if (h->flush_sink) cudaFree(h->flush_sink);
    if (h->gather_host) cudaFreeHost(h->gather_host);
    for (void *p : h->allocs) cudaFreeAsync(p, h->stream);
    if (h->stream) cudaStreamSynchronize(h->stream);
    if (h->ctrl_host) pin_release(h->ctrl_host);        // res_host lives in the same pinned block
    for (auto &e : h->ev) if (e) cudaEventDestroy(e);
    for (auto &e : h->kev) if (e) cudaEventDestroy(e);
    if (h->span0) cudaEventDestroy(h->span0);
    if (h->span1) cudaEventDestroy(h->span1);
    if (h->stream) cudaStreamDestroy(h->stream);
    delete h;
}

int64_t ea_nvar(const ea_handle_t *h) { return h ? h->nvar : 0; }

int ea_init_solution(ea_handle_t *h, double rho_pq, double rho_va) {
    if (!h) return EA_ERR_ARG;
    CK(cudaSetDevice(h->device));
    for (int f = 0; f < EA_NUM_FIELDS; ++f) CK(cudaMemsetAsync(h->fields[f], 0, sizeof(double) * (size_t)h->nint, h->stream));
    h->zsel = 0;
    const int n = std::max(h->gpad, (int)h->nline);
    k_init_solution<<<nblocks(n, 128), 128, 0, h->stream>>>(h->d, rho_pq, rho_va);
    CK(cudaGetLastError());
    CK(cudaStreamSynchronize(h->stream));
    return EA_OK;
}

// Sum of one scalar over the ranks of a partitioned handle (rank order, identical on every rank).
static int allgather_scalar_sum(ea_handle *h, double local, double *total) {
    if (!h->d.partitioned) { *total = local; return EA_OK; }
    if (!h->comm) return fail(h, EA_ERR_STATE, "partitioned handle without a communicator (call ea_comm_init)");
    CK(cudaMemcpyAsync(h->d.sendbuf, &local, sizeof(double), cudaMemcpyHostToDevice, h->stream));
    ncclResult_t r = g_nccl.AllGather(h->d.sendbuf, h->gather_dev, (size_t)h->d.stride, ncclDouble, h->comm, h->stream);
    if (r != ncclSuccess) return fail(h, EA_ERR_NCCL, "ncclAllGather: %s", g_nccl.GetErrorString(r));
    CK(cudaMemcpyAsync(h->gather_host, h->gather_dev, sizeof(double) * (size_t)h->d.stride * h->part_nranks, cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    double s = 0.0;
    for (int r2 = 0; r2 < h->part_nranks; ++r2) s += h->gather_host[(size_t)r2 * h->d.stride];
    *total = s;
    return EA_OK;
}

static int norm_of(ea_handle *h, const double *x, double *out) {
    const int n = (int)h->n_owned_entries;            // the whole vector, or this rank's owned prefix
    const int grid = std::min(h->max_blocks, nblocks(n, RBLOCK));
    k_norm<<<grid, RBLOCK, 0, h->stream>>>(n, x, h->d.partials, &h->d.ctrl->ticket, h->res_dev);
    CK(cudaGetLastError());
    h->n_other++;
    CK(cudaMemcpyAsync(h->res_host, h->res_dev, sizeof(double), cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    if (!h->d.partitioned) { *out = h->res_host[0]; return EA_OK; }
    double tot = 0.0;
    int rc = allgather_scalar_sum(h, h->res_host[0] * h->res_host[0], &tot);
    if (rc) return rc;
    *out = std::sqrt(tot);
    return EA_OK;
}

int ea_outer_prestep(ea_handle_t *h, double *norm_z_prev) {
    if (!h || !norm_z_prev) return EA_ERR_ARG;
    CK(cudaSetDevice(h->device));
    return norm_of(h, h->d.zbuf[h->zsel], norm_z_prev);
}

int ea_inner_prestep(ea_handle_t *h) {
    if (!h) return EA_ERR_ARG;
    CK(cudaSetDevice(h->device));
    CK(cudaMemcpyAsync(h->d.zbuf[h->zsel ^ 1], h->d.zbuf[h->zsel], sizeof(double) * (size_t)h->nint, cudaMemcpyDeviceToDevice, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    return EA_OK;
}

static int timed_x(ea_handle *h, int64_t inner, int32_t max_auglag, double mu_max, double scale, int lines, int gens) {
    if (inner < 1) return fail(h, EA_ERR_ARG, "update_x: info.inner must be >= 1");
    CK(cudaSetDevice(h->device));
    CK(cudaEventRecord(h->ev[0], h->stream));
    int rc = launch_x(h, inner, h->zsel, max_auglag, mu_max, scale, lines, gens);
    if (rc) return rc;
    CK(cudaEventRecord(h->ev[1], h->stream));
    CK(cudaStreamSynchronize(h->stream));
    const double t = elapsed_s(h->ev[0], h->ev[1]);
    h->tm.x += t;
    if (lines) h->tm.line += t; else h->tm.gen += t;
    return EA_OK;
}

int ea_update_x_gen(ea_handle_t *h) { return h ? timed_x(h, 1, 1, 1.0, 1.0, 0, 1) : EA_ERR_ARG; }

int ea_update_x_line(ea_handle_t *h, int64_t inner, int32_t max_auglag, double mu_max, double scale) {
    return h ? timed_x(h, inner, max_auglag, mu_max, scale, 1, 0) : EA_ERR_ARG;
}

int ea_update_x(ea_handle_t *h, int64_t inner, int32_t max_auglag, double mu_max, double scale) {
    if (!h) return EA_ERR_ARG;
    int rc = ea_update_x_gen(h);
    if (rc) return rc;
    return ea_update_x_line(h, inner, max_auglag, mu_max, scale);
}

int ea_update_xbar(ea_handle_t *h) {
    if (!h) return EA_ERR_ARG;
    CK(cudaSetDevice(h->device));
    CK(cudaEventRecord(h->ev[0], h->stream));
    k_bus<false><<<nblocks((int64_t)h->d.n_bus_warps * 32, BBLOCK), BBLOCK, 0, h->stream>>>(h->d, h->zsel, 0.0);
    CK(cudaGetLastError());
    CK(cudaEventRecord(h->ev[1], h->stream));
    CK(cudaStreamSynchronize(h->stream));
    h->tm.bus += elapsed_s(h->ev[0], h->ev[1]);
    return EA_OK;
}

int ea_update_z(ea_handle_t *h, double beta) {
    if (!h) return EA_ERR_ARG;
    CK(cudaSetDevice(h->device));
    CK(cudaEventRecord(h->ev[0], h->stream));
    k_update_z<<<nblocks(h->nint, 256), 256, 0, h->stream>>>(h->nint, h->d.zbuf[h->zsel], h->d.lz, h->d.l, h->d.rho, h->d.u, h->d.v, beta);
    CK(cudaGetLastError());
    CK(cudaEventRecord(h->ev[1], h->stream));
    CK(cudaStreamSynchronize(h->stream));
    h->tm.z += elapsed_s(h->ev[0], h->ev[1]);
    return EA_OK;
}

int ea_update_l(ea_handle_t *h, double beta) {
    if (!h) return EA_ERR_ARG;
    CK(cudaSetDevice(h->device));
    CK(cudaEventRecord(h->ev[0], h->stream));
    k_update_l<<<nblocks(h->nint, 256), 256, 0, h->stream>>>(h->nint, h->d.l, h->d.lz, h->d.zbuf[h->zsel], beta);
    CK(cudaGetLastError());
    CK(cudaEventRecord(h->ev[1], h->stream));
    CK(cudaStreamSynchronize(h->stream));
    h->tm.l += elapsed_s(h->ev[0], h->ev[1]);
    return EA_OK;
}

int ea_update_lz(ea_handle_t *h, double beta, double max_multiplier) {
    if (!h) return EA_ERR_ARG;
    CK(cudaSetDevice(h->device));
    CK(cudaEventRecord(h->ev[0], h->stream));
    k_update_lz<<<nblocks(h->nint, 256), 256, 0, h->stream>>>(h->nint, h->d.lz, h->d.zbuf[h->zsel], beta, max_multiplier);
    CK(cudaGetLastError());
    h->n_other++;
    CK(cudaEventRecord(h->ev[1], h->stream));
    CK(cudaStreamSynchronize(h->stream));
    h->tm.lz += elapsed_s(h->ev[0], h->ev[1]);
    return EA_OK;
}

static int residual_vectors(ea_handle *h, double out[4]) {
    const int grid = std::min(h->max_blocks, nblocks(h->nint, RBLOCK));
    k_residual<<<grid, RBLOCK, 0, h->stream>>>(h->nint, h->d.u, h->d.v, h->d.zbuf[h->zsel], h->d.zbuf[h->zsel ^ 1],
                                               h->d.rp, h->d.rd, h->d.axby, h->d.partials, &h->d.ctrl->ticket, h->res_dev);
    CK(cudaGetLastError());
    h->n_other++;
    CK(cudaMemcpyAsync(h->res_host, h->res_dev, 4 * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    if (out) for (int k = 0; k < 4; ++k) out[k] = h->res_host[k];
    return EA_OK;
}

int ea_update_residual(ea_handle_t *h, double out[4]) {
    if (!h || !out) return EA_ERR_ARG;
    if (h->d.partitioned)
        return fail(h, EA_ERR_STATE, "step-wise operators act on the local sub-grid only; a partitioned handle is driven "
                                     "through ea_inner_iteration / ea_run_inner / ea_admm_two_level");
    CK(cudaSetDevice(h->device));
    return residual_vectors(h, out);
}

int ea_poststep(ea_handle_t *h, double *objval) {
    if (!h || !objval) return EA_ERR_ARG;
    CK(cudaSetDevice(h->device));
    std::vector<double> ug((size_t)std::max(h->gpad, 1));
    CK(cudaMemcpyAsync(ug.data(), h->d.u, sizeof(double) * (size_t)h->gpad, cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    // acopf_admm_prepoststep_gpu.jl:35-41: host-side sum over generators in reference order, unscaled cost
    std::vector<double> pg((size_t)h->ngen);
    for (int k = 0; k < (int)h->ngen; ++k) pg[h->gen_of_slot[k]] = ug[2 * k];
    double obj = 0.0;
    for (int g = 0; g < (int)h->ngen; ++g) {
        const double p = h->d.baseMVA * pg[g];
        obj += h->c2[g] * (p * p) + h->c1[g] * p + h->c0[g];
    }
    return allgather_scalar_sum(h, obj, objval);
}

// ---- fused path --------------------------------------------------------------------------
static int enqueue_iteration(ea_handle *h, int max_auglag, double mu_max, double scale, int slot) {
    cudaEvent_t *e = nullptr;
    if (h->kernel_timing) {
        while ((int)h->kev.size() < 3 * (slot + 1)) {
            cudaEvent_t ev;
            CK(cudaEventCreate(&ev));
            h->kev.push_back(ev);
        }
        e = &h->kev[3 * slot];
        // measurement mode: evict the L2 (a write larger than it) before the iteration, outside the event brackets
        if (h->flush_buf) {
            CK(cudaMemsetAsync(h->flush_buf, slot & 0xff, h->flush_bytes, h->stream));
            // "l2_flush_clean": the write leaves the L2 full of DIRTY lines, whose write-back the timed kernel's misses
            // would then wait for; a read of a second buffer of the same size replaces them with clean lines
            if (h->flush_clean)
                k_l2_read<<<1184, 256, 0, h->stream>>>(reinterpret_cast<const double4 *>((char *)h->flush_buf + h->flush_bytes),
                                                       h->flush_bytes / sizeof(double4), h->flush_sink);
        }
        CK(cudaEventRecord(e[0], h->stream));
    }
    int rc = launch_x(h, 0, 0, max_auglag, mu_max, scale, 1, 1);
    if (rc) return rc;
    if (e && h->kernel_timing == 1) CK(cudaEventRecord(e[1], h->stream));      // (2: the iteration as a whole)
    k_bus<true><<<nblocks((int64_t)h->d.n_bus_warps * 32, BBLOCK), BBLOCK, 0, h->stream>>>(h->d, -1, 0.0);
    CK(cudaGetLastError());
    h->n_bus++;
    if (h->d.partitioned && !h->loopback) {
        // the one exchange of the iteration: xbar halves of the cut-branch ends + residual partial sums
        if (!h->d.peer_mode) {
            if (!h->comm) return fail(h, EA_ERR_STATE, "partitioned handle without a communicator (call ea_comm_init)");
            ncclResult_t r = g_nccl.AllGather(h->d.sendbuf, h->gather_dev, (size_t)h->d.stride, ncclDouble, h->comm, h->stream);
            if (r != ncclSuccess) return fail(h, EA_ERR_NCCL, "ncclAllGather: %s", g_nccl.GetErrorString(r));
        }   // peer mode: the bus kernel has already stored the segment into the peers' buffers
        k_finish<<<1, FBLOCK, 0, h->stream>>>(h->d);
        CK(cudaGetLastError());
        h->n_other++;
    }
    if (e) CK(cudaEventRecord(e[2], h->stream));
    return EA_OK;
}

static int fetch_ctrl(ea_handle *h) {
    CK(cudaMemcpyAsync(h->ctrl_host, h->d.ctrl, sizeof(Ctrl), cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    h->zsel = h->ctrl_host->zsel;
    return EA_OK;
}

int ea_inner_iteration(ea_handle_t *h, int64_t inner, double beta, int32_t max_auglag, double mu_max, double scale,
                       double out[4]) {
    if (!h || !out) return EA_ERR_ARG;
    if (inner < 1) return fail(h, EA_ERR_ARG, "ea_inner_iteration: info.inner must be >= 1");
    CK(cudaSetDevice(h->device));
    int rc = sync_ctrl_to_device(h, beta, -1.0, inner - 1, std::numeric_limits<long long>::max());
    if (rc) return rc;
    const int kt = h->kernel_timing; h->kernel_timing = 0;
    rc = enqueue_iteration(h, max_auglag, mu_max, scale, 0);
    h->kernel_timing = kt;
    if (rc) return rc;
    if ((rc = fetch_ctrl(h))) return rc;
    for (int k = 0; k < 4; ++k) out[k] = h->ctrl_host->res[k];
    return EA_OK;
}

// `chunk` iterations of the fused loop (2 launches each) captured once and replayed: launch-bound grids (a few thousand
// branches) spend a third of the iteration between kernels otherwise. Launches after the loop has finished are no-ops
// on the device (ctrl->done), so a replay may overshoot the iteration limit.
static int build_loop_graph(ea_handle *h, int chunk, int max_auglag, double mu_max, double scale) {
    if (h->graph && h->graph_chunk == chunk && h->graph_max_auglag == max_auglag && h->graph_mu_max == mu_max &&
        h->graph_scale == scale && h->graph_count_work == h->d.count_work) return EA_OK;
    drop_loop_graph(h);
    build_pow_table(h, mu_max);
    const long long n_x = h->n_x, n_bus = h->n_bus, n_other = h->n_other;
    cudaGraph_t g = nullptr;
    CK(cudaStreamBeginCapture(h->stream, cudaStreamCaptureModeThreadLocal));
    int rc = EA_OK;
    for (int i = 0; i < chunk && rc == EA_OK; ++i) rc = enqueue_iteration(h, max_auglag, mu_max, scale, i);
    const cudaError_t e = cudaStreamEndCapture(h->stream, &g);
    h->n_x = n_x; h->n_bus = n_bus; h->n_other = n_other;             // counted per replay instead
    if (rc) { if (g) cudaGraphDestroy(g); return rc; }
    if (e != cudaSuccess) return fail(h, EA_ERR_CUDA, "cudaStreamEndCapture failed: %s", cudaGetErrorString(e));
    const cudaError_t e2 = cudaGraphInstantiate(&h->graph, g, 0);
    cudaGraphDestroy(g);
    if (e2 != cudaSuccess) { h->graph = nullptr; return fail(h, EA_ERR_CUDA, "cudaGraphInstantiate failed: %s", cudaGetErrorString(e2)); }
    h->graph_chunk = chunk; h->graph_max_auglag = max_auglag; h->graph_mu_max = mu_max; h->graph_scale = scale;
    h->graph_count_work = h->d.count_work;
    // move the executable graph to the device now: otherwise the first replay pays for it (~0.5 ms with 32 nodes)
    CK(cudaGraphUpload(h->graph, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    return EA_OK;
}

int ea_run_inner_from(ea_handle_t *h, int64_t outer, double beta, int64_t inner_start, int64_t inner_limit,
                      int32_t max_auglag, double mu_max, double scale, int32_t chunk, int64_t *inner_done, double out[4]) {
    if (!h || !inner_done || !out) return EA_ERR_ARG;
    if (outer < 1 || inner_start < 0 || inner_limit < inner_start) return fail(h, EA_ERR_ARG, "ea_run_inner: bad outer / inner range");
    CK(cudaSetDevice(h->device));
    if (chunk <= 0) chunk = h->default_chunk;
    const double eps_pri = std::sqrt((double)h->nvar_global) / (2500.0 * (double)outer);    // admm_two_level.jl:45
    if (inner_limit == inner_start) { *inner_done = inner_start; for (int k = 0; k < 4; ++k) out[k] = 0.0; return EA_OK; }
    int rc;
    int64_t enq = inner_start;
    // (partitioned handles too: the exchange is part of the chunk - peer stores inside the bus kernel, or the captured
    //  ncclAllGather; every rank replays the same graph, launches after `done` are no-ops on every rank alike)
    const bool graph = h->use_graph && !h->kernel_timing && !h->loopback && chunk > 1 && inner_limit - inner_start >= chunk;
    if (graph && (rc = build_loop_graph(h, chunk, max_auglag, mu_max, scale))) return rc;     // host work: before the span
    CK(cudaEventRecord(h->span0, h->stream));
    if ((rc = sync_ctrl_to_device(h, beta, eps_pri, inner_start, inner_limit))) return rc;
    for (;;) {
        const int64_t todo = graph ? chunk : std::min<int64_t>(chunk, inner_limit - enq);
        if (graph) {
            CK(cudaGraphLaunch(h->graph, h->stream));
            h->n_replay++; h->n_x += chunk; h->n_bus += chunk;
            if (h->d.partitioned) h->n_other += chunk;                 // k_finish
        } else {
            for (int64_t i = 0; i < todo; ++i)
                if ((rc = enqueue_iteration(h, max_auglag, mu_max, scale, (int)i))) return rc;
        }
        enq += todo;
        if ((rc = fetch_ctrl(h))) return rc;
        if (h->kernel_timing) {
            // iterations that really ran in this chunk (later launches were no-ops)
            const int64_t ran = std::min<int64_t>(todo, h->ctrl_host->inner - (enq - todo));
            h->n_timed += std::max<int64_t>(ran, 0);
            for (int64_t i = 0; i < ran; ++i) {
                if (h->kernel_timing == 1) {
                    h->t_x += elapsed_s(h->kev[3 * i], h->kev[3 * i + 1]);
                    h->t_bus += elapsed_s(h->kev[3 * i + 1], h->kev[3 * i + 2]);
                } else h->t_x += elapsed_s(h->kev[3 * i], h->kev[3 * i + 2]);       // both kernels, back to back
            }
        }
        if (h->ctrl_host->done || enq >= inner_limit) break;
    }
    *inner_done = h->ctrl_host->inner;
    for (int k = 0; k < 4; ++k) out[k] = h->ctrl_host->res[k];
    // leave rp / rd / Ax_plus_By as the reference's last admm_update_residual would
    rc = residual_vectors(h, nullptr);
    if (rc) return rc;
    CK(cudaEventRecord(h->span1, h->stream));
    CK(cudaEventSynchronize(h->span1));
    h->span_s += elapsed_s(h->span0, h->span1);
    return EA_OK;
}

int ea_run_inner(ea_handle_t *h, int64_t outer, double beta, int64_t inner_iterlim, int32_t max_auglag, double mu_max,
                 double scale, int32_t chunk, int64_t *inner_done, double out[4]) {
    if (inner_iterlim < 0) return h ? fail(h, EA_ERR_ARG, "ea_run_inner: bad inner_iterlim") : EA_ERR_ARG;
    return ea_run_inner_from(h, outer, beta, 0, inner_iterlim, max_auglag, mu_max, scale, chunk, inner_done, out);
}

int ea_get_kernel_times(ea_handle_t *h, double out[8]) {
    if (!h || !out) return EA_ERR_ARG;
    // with kernel_timing on, [1] and [3] count the iterations the duration sums cover; the launches that found
    // `done` set and returned at once (tail of a chunk) are reported with the other launches in [5]
    const bool kt = h->kernel_timing && h->n_timed > 0;
    out[0] = h->span_s; out[1] = kt ? (double)h->n_timed : (double)h->n_x; out[2] = h->t_x;
    out[3] = kt ? (double)h->n_timed : (double)h->n_bus; out[4] = h->t_bus;
    out[5] = (double)h->n_other + (kt ? (double)(h->n_x + h->n_bus - 2 * h->n_timed) : 0.0); out[6] = 0.0; out[7] = 0.0;
    if (h->d.count_work > 1) {      // diagnostics: x-update phase split of the LAST launch since reset (seconds)
        Counters c;
        cudaStreamSynchronize(h->stream);
        if (cudaMemcpy(&c, h->d.counters, sizeof(c), cudaMemcpyDeviceToHost) == cudaSuccess && c.t[2] > c.t[0]) {
            out[6] = 1e-9 * (double)(c.t[1] - c.t[0]);     // start -> work queue empty
            out[7] = 1e-9 * (double)(c.t[2] - c.t[0]);     // start -> last warp done
        }
    }
    return EA_OK;
}

int ea_admm_two_level(ea_handle_t *h, const ea_params_t *par, ea_info_t *info) {
    if (!h || !par || !info) return EA_ERR_ARG;
    CK(cudaSetDevice(h->device));
    const double sqrt_d = std::sqrt((double)h->nvar_global);
    const double OUTER_TOL = sqrt_d * par->outer_eps;
    memset(info, 0, sizeof(*info));
    info->mismatch = INFINITY; info->norm_z_prev = INFINITY; info->norm_z_curr = INFINITY;
    double beta = par->initial_beta;
    double res[4];
    int rc;
    if (par->verbose > 0 && !h->d.partitioned) {     // admm_two_level.jl:15-25
        if ((rc = ea_update_residual(h, res))) return rc;
        info->primres = res[0]; info->dualres = res[1]; info->norm_z_curr = res[2]; info->mismatch = res[3];
        printf("%8s  %8s  %10s  %10s  %10s  %10s  %10s  %10s  %10s  %10s  %10s\n", "Outer", "Inner", "Objval", "AugLag",
               "PrimRes", "EpsPrimRes", "DualRes", "||z||", "||Ax+By||", "OuterTol", "Beta");
        printf("%8lld  %8lld  %10.3e  %10.3e  %10.3e  %10.3e  %10.3e  %10.3e  %10.3e  %10.3e  %10.3e\n", 0ll, 0ll, 0.0, 0.0,
               info->primres, info->eps_pri, info->dualres, info->norm_z_curr, info->mismatch, OUTER_TOL, beta);
    }
    info->status = EA_STATUS_ITERATION_LIMIT;
    h->tm = Timers();
    const auto t0 = std::chrono::steady_clock::now();
    while (info->outer < par->outer_iterlim) {
        info->outer++;
        if ((rc = ea_outer_prestep(h, &info->norm_z_prev))) return rc;
        info->inner = 0;
        if (par->verbose > 0) {
            // per-iteration table: one host round trip per inner iteration
            const bool talk = !h->d.partitioned || h->part_rank == 0;
            while (info->inner < par->inner_iterlim) {
                info->inner++; info->cumul++;
                if ((rc = ea_inner_iteration(h, info->inner, beta, par->max_auglag, par->mu_max, par->scale, res))) return rc;
                info->primres = res[0]; info->dualres = res[1]; info->norm_z_curr = res[2]; info->mismatch = res[3];
                info->eps_pri = sqrt_d / (2500.0 * (double)info->outer);
                if (talk && (info->cumul % 50) == 0)
                    printf("%8s  %8s  %10s  %10s  %10s  %10s  %10s  %10s  %10s  %10s  %10s\n", "Outer", "Inner", "Objval",
                           "AugLag", "PrimRes", "EpsPrimRes", "DualRes", "||z||", "||Ax+By||", "OuterTol", "Beta");
                if (talk) printf("%8lld  %8lld  %10.3e  %10.3e  %10.3e  %10.3e  %10.3e  %10.3e  %10.3e  %10.3e  %10.3e\n",
                       (long long)info->outer, (long long)info->inner, info->objval, info->auglag, info->primres,
                       info->eps_pri, info->dualres, info->norm_z_curr, info->mismatch, OUTER_TOL, beta);
                if (info->primres <= info->eps_pri) break;
            }
            if ((rc = residual_vectors(h, nullptr))) return rc;
        } else {
            int64_t done = 0;
            if ((rc = ea_run_inner(h, info->outer, beta, par->inner_iterlim, par->max_auglag, par->mu_max, par->scale, 0,
                                   &done, res))) return rc;
            info->inner = done; info->cumul += done;
            if (done > 0) { info->primres = res[0]; info->dualres = res[1]; info->norm_z_curr = res[2]; info->mismatch = res[3]; }
            info->eps_pri = sqrt_d / (2500.0 * (double)info->outer);
        }
        if (info->mismatch <= OUTER_TOL) { info->status = EA_STATUS_SOLVED; break; }
        if ((rc = ea_update_lz(h, beta, par->MAX_MULTIPLIER))) return rc;
        if (info->norm_z_curr > par->theta * info->norm_z_prev) beta = std::min(par->inc_c * beta, 1e24);
    }
    info->time_overall = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    info->beta = beta;
    info->time_x_update = h->tm.x; info->time_xbar_update = h->tm.bus; info->time_z_update = h->tm.z;
    info->time_l_update = h->tm.l; info->time_lz_update = h->tm.lz;
    info->time_generators = h->tm.gen; info->time_branches = h->tm.line; info->time_buses = h->tm.bus;
    return ea_poststep(h, &info->objval);
}

// ---- data access ----------------------------------------------------------------------------
int ea_get_vector(ea_handle_t *h, int field, double *host, int64_t n) {
    if (!h || !host) return EA_ERR_ARG;
    if (field < 0 || field >= EA_NUM_FIELDS) return fail(h, EA_ERR_ARG, "ea_get_vector: bad field %d", field);
    if (n != h->nvar) return fail(h, EA_ERR_ARG, "ea_get_vector: n = %lld, expected nvar = %lld", (long long)n, (long long)h->nvar);
    CK(cudaSetDevice(h->device));
    if (n == 0) return EA_OK;
    k_gather<<<nblocks(n, 256), 256, 0, h->stream>>>((int)n, h->ref2int, field_ptr(h, field), h->staging);
    CK(cudaGetLastError());
    CK(cudaMemcpyAsync(host, h->staging, sizeof(double) * (size_t)n, cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    return EA_OK;
}

int ea_set_vector(ea_handle_t *h, int field, const double *host, int64_t n) {
    if (!h || !host) return EA_ERR_ARG;
    if (field < 0 || field >= EA_NUM_FIELDS) return fail(h, EA_ERR_ARG, "ea_set_vector: bad field %d", field);
    if (n != h->nvar) return fail(h, EA_ERR_ARG, "ea_set_vector: n = %lld, expected nvar = %lld", (long long)n, (long long)h->nvar);
    CK(cudaSetDevice(h->device));
    if (n == 0) return EA_OK;
    CK(cudaMemcpyAsync(h->staging, host, sizeof(double) * (size_t)n, cudaMemcpyHostToDevice, h->stream));
    k_scatter<<<nblocks(n, 256), 256, 0, h->stream>>>((int)n, h->ref2int, h->staging, field_ptr(h, field));
    CK(cudaGetLastError());
    CK(cudaStreamSynchronize(h->stream));
    return EA_OK;
}

int ea_get_membuf(ea_handle_t *h, int row, double *host, int64_t n) {
    if (!h || !host) return EA_ERR_ARG;
    if (row < 1 || row > 31) return fail(h, EA_ERR_ARG, "ea_get_membuf: row %d outside 1..31", row);
    if (n != h->nline) return fail(h, EA_ERR_ARG, "ea_get_membuf: n = %lld, expected nline = %lld", (long long)n, (long long)h->nline);
    CK(cudaSetDevice(h->device));
    if (n == 0) return EA_OK;
    const double *src = nullptr;
    if (row >= 25 && row <= 27) src = h->d.als + (size_t)(row - 25) * h->nline;
    else if (row == 29) src = h->d.rateA;
    else if (row <= 24) {
        k_membuf_row<<<nblocks(n, 256), 256, 0, h->stream>>>(h->d, h->zsel, row - 1, h->staging);
        CK(cudaGetLastError());
        src = h->staging;
    } else { memset(host, 0, sizeof(double) * (size_t)n); return EA_OK; }
    CK(cudaMemcpyAsync(host, src, sizeof(double) * (size_t)n, cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    return EA_OK;
}

int ea_set_membuf(ea_handle_t *h, int row, const double *host, int64_t n) {
    if (!h || !host) return EA_ERR_ARG;
    if (n != h->nline) return fail(h, EA_ERR_ARG, "ea_set_membuf: n = %lld, expected nline = %lld", (long long)n, (long long)h->nline);
    CK(cudaSetDevice(h->device));
    double *dst = nullptr;
    if (row >= 25 && row <= 27) dst = h->d.als + (size_t)(row - 25) * h->nline;
    else if (row == 29) dst = const_cast<double *>(h->d.rateA);
    else return fail(h, EA_ERR_ARG, "ea_set_membuf: only rows 25, 26, 27 and 29 hold state (got %d)", row);
    if (n == 0) return EA_OK;
    CK(cudaMemcpyAsync(dst, host, sizeof(double) * (size_t)n, cudaMemcpyHostToDevice, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    return EA_OK;
}

int ea_set_load(ea_handle_t *h, const double *Pd, const double *Qd, int64_t nbus) {
    if (!h || !Pd || !Qd) return EA_ERR_ARG;
    if (nbus != h->nbus) return fail(h, EA_ERR_ARG, "ea_set_load: nbus = %lld, expected %lld", (long long)nbus, (long long)h->nbus);
    CK(cudaSetDevice(h->device));
    std::vector<double> pd((size_t)nbus), qd((size_t)nbus);
    for (int64_t b = 0; b < nbus; ++b) { pd[b] = Pd[b] / h->d.baseMVA; qd[b] = Qd[b] / h->d.baseMVA; }
    CK(cudaMemcpy(const_cast<double *>(h->d.pd_pu), pd.data(), sizeof(double) * (size_t)nbus, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(const_cast<double *>(h->d.qd_pu), qd.data(), sizeof(double) * (size_t)nbus, cudaMemcpyHostToDevice));
    return EA_OK;
}

int ea_set_pg_bounds(ea_handle_t *h, const double *lo, const double *hi, int64_t ngen) {
    if (!h || !lo || !hi) return EA_ERR_ARG;
    if (ngen != h->ngen) return fail(h, EA_ERR_ARG, "ea_set_pg_bounds: ngen = %lld, expected %lld", (long long)ngen, (long long)h->ngen);
    CK(cudaSetDevice(h->device));
    std::vector<double> a((size_t)ngen), b((size_t)ngen);
    for (int64_t k = 0; k < ngen; ++k) { a[k] = lo[h->gen_of_slot[k]]; b[k] = hi[h->gen_of_slot[k]]; }
    CK(cudaMemcpy(const_cast<double *>(h->d.pgmin_curr), a.data(), sizeof(double) * (size_t)ngen, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(const_cast<double *>(h->d.pgmax_curr), b.data(), sizeof(double) * (size_t)ngen, cudaMemcpyHostToDevice));
    return EA_OK;
}

int ea_get_counters(ea_handle_t *h, ea_counters_t *out) {
    if (!h || !out) return EA_ERR_ARG;
    CK(cudaSetDevice(h->device));
    Counters c;
    CK(cudaStreamSynchronize(h->stream));
    CK(cudaMemcpy(&c, h->d.counters, sizeof(c), cudaMemcpyDeviceToHost));
    out->line_calls = (int64_t)c.v[0]; out->auglag_iters = (int64_t)c.v[1]; out->tron_evals = (int64_t)c.v[2];
    out->cg_iters = (int64_t)c.v[3]; out->chol_shifts = (int64_t)c.v[4]; out->rejected_steps = (int64_t)c.v[5];
    out->max_auglag_hits = (int64_t)c.v[6]; out->max_evals_lane = (int64_t)c.v[7];
    return EA_OK;
}

int ea_reset_counters(ea_handle_t *h) {
    if (!h) return EA_ERR_ARG;
    CK(cudaSetDevice(h->device));
    CK(cudaStreamSynchronize(h->stream));
    CK(cudaMemset(h->d.counters, 0, sizeof(Counters)));
    { unsigned long long big = ~0ull; CK(cudaMemcpy(&h->d.counters->t[0], &big, 8, cudaMemcpyHostToDevice)); CK(cudaMemcpy(&h->d.counters->t[1], &big, 8, cudaMemcpyHostToDevice)); }
    h->span_s = 0.0; h->n_x = h->n_bus = h->n_other = 0; h->t_x = h->t_bus = 0.0; h->n_timed = 0;
    return EA_OK;
}

int ea_set_option(ea_handle_t *h, const char *name, double value) {
    if (!h || !name) return EA_ERR_ARG;
    if (!strcmp(name, "count_work")) { h->d.count_work = (int)value; return EA_OK; }   // 2: also phase timestamps
    if (!strcmp(name, "chunk")) { h->default_chunk = std::max(1, (int)value); return EA_OK; }
    if (!strcmp(name, "kernel_timing")) { h->kernel_timing = (value == 2.0) ? 2 : (value != 0.0); return EA_OK; }
    if (!strcmp(name, "use_graph")) { h->use_graph = value != 0.0; return EA_OK; }
    if (!strcmp(name, "l2_flush_mb")) {             // > 0: with kernel_timing, write this many MB before every iteration
        CK(cudaSetDevice(h->device));
        CK(cudaStreamSynchronize(h->stream));
        if (h->flush_buf) { cudaFree(h->flush_buf); h->flush_buf = nullptr; h->flush_bytes = 0; }
        if (value > 0.0) {
            h->flush_bytes = (size_t)(value * 1048576.0);
            if (!h->flush_sink) CK(cudaMalloc(&h->flush_sink, sizeof(double)));
            if (cudaMalloc(&h->flush_buf, 2 * h->flush_bytes) != cudaSuccess ||
                cudaMemset(h->flush_buf, 0, 2 * h->flush_bytes) != cudaSuccess) {
                h->flush_buf = nullptr; h->flush_bytes = 0;
                return fail(h, EA_ERR_ALLOC, "ea_set_option: cannot allocate the L2 flush buffer");
            }
        }
        return EA_OK;
    }
    if (!strcmp(name, "l2_flush_clean")) { h->flush_clean = value != 0.0; return EA_OK; }
    return fail(h, EA_ERR_ARG, "ea_set_option: unknown option '%s'", name);
}

// ---- bus-partitioned multi-GPU mode ------------------------------------------------------------
int ea_set_partition(ea_handle_t *h, int32_t rank, int32_t nranks, int64_t n_owned_bus, int64_t n_send,
                     const int64_t *send_line, const int64_t *send_end, int64_t n_ghost, const int64_t *ghost_line,
                     const int64_t *ghost_end, const int64_t *ghost_src_rank, const int64_t *ghost_src_pos, int64_t max_send,
                     int64_t nvar_global) {
    if (!h) return EA_ERR_ARG;
    if (nranks < 1 || rank < 0 || rank >= nranks || n_owned_bus < 0 || n_owned_bus > h->nbus || n_send < 0 || n_ghost < 0 ||
        max_send < n_send || nvar_global < h->nvar - 8 * n_ghost / 2 || (n_send && (!send_line || !send_end)) ||
        (n_ghost && (!ghost_line || !ghost_end || !ghost_src_rank || !ghost_src_pos)))
        return fail(h, EA_ERR_ARG, "ea_set_partition: bad argument");
    if (h->d.partitioned) return fail(h, EA_ERR_STATE, "ea_set_partition: already partitioned");
    CK(cudaSetDevice(h->device));
    // half-slot of (line, end) from the layout built in ea_create
    std::vector<int> slot_from((size_t)h->nline), slot_to((size_t)h->nline), hstart((size_t)h->nbus + 1);
    CK(cudaMemcpy(slot_from.data(), h->d.slot_from, sizeof(int) * (size_t)h->nline, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(slot_to.data(), h->d.slot_to, sizeof(int) * (size_t)h->nline, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(hstart.data(), h->d.hstart, sizeof(int) * ((size_t)h->nbus + 1), cudaMemcpyDeviceToHost));
    const int owned_halves = hstart[(size_t)n_owned_bus];
    const int stride = 4 + 4 * (int)max_send;
    std::vector<int> send_pos(2 * (size_t)h->nline, -1), ghost_slot((size_t)n_ghost), ghost_src((size_t)n_ghost);
    for (int64_t k = 0; k < n_send; ++k) {
        if (send_line[k] < 0 || send_line[k] >= h->nline || (send_end[k] != 0 && send_end[k] != 1))
            return fail(h, EA_ERR_ARG, "ea_set_partition: bad send entry %lld", (long long)k);
        const int s = send_end[k] ? slot_to[send_line[k]] : slot_from[send_line[k]];
        if (s >= owned_halves) return fail(h, EA_ERR_ARG, "ea_set_partition: send entry %lld is not an owned branch end", (long long)k);
        send_pos[s] = (int)k;
    }
    for (int64_t k = 0; k < n_ghost; ++k) {
        if (ghost_line[k] < 0 || ghost_line[k] >= h->nline || (ghost_end[k] != 0 && ghost_end[k] != 1) ||
            ghost_src_rank[k] < 0 || ghost_src_rank[k] >= nranks || ghost_src_rank[k] == rank || ghost_src_pos[k] < 0 ||
            ghost_src_pos[k] >= max_send)
            return fail(h, EA_ERR_ARG, "ea_set_partition: bad ghost entry %lld", (long long)k);
        const int s = ghost_end[k] ? slot_to[ghost_line[k]] : slot_from[ghost_line[k]];
        if (s < owned_halves) return fail(h, EA_ERR_ARG, "ea_set_partition: ghost entry %lld is an owned branch end", (long long)k);
        ghost_slot[k] = s;
        ghost_src[k] = (int)(ghost_src_rank[k] * stride + 4 + 4 * ghost_src_pos[k]);
    }
    if ((int64_t)owned_halves + n_ghost != 2 * h->nline)
        return fail(h, EA_ERR_ARG, "ea_set_partition: owned + ghost branch ends do not cover the local grid");
    int rc;
    Dev &d = h->d;
    if ((rc = dev_upload(h, const_cast<int **>(&d.send_pos), send_pos))) return rc;
    if ((rc = dev_upload(h, const_cast<int **>(&d.ghost_slot), ghost_slot))) return rc;
    if ((rc = dev_upload(h, const_cast<int **>(&d.ghost_src), ghost_src))) return rc;
    if ((rc = dev_alloc(h, &h->gather_dev, (size_t)stride * nranks))) return rc;
    if (cudaMallocHost((void **)&h->gather_host, sizeof(double) * (size_t)stride * nranks) != cudaSuccess)
        return fail(h, EA_ERR_ALLOC, "cudaMallocHost failed");
    d.gather = h->gather_dev;
    d.sendbuf = h->gather_dev + (size_t)rank * stride;
    d.partitioned = 1; d.rank = rank; d.nranks = nranks; d.n_ghost = (int)n_ghost; d.stride = stride;
    d.nbus_active = (int)n_owned_bus;
    if ((rc = build_bus_warps(h, hstart, (int)n_owned_bus))) return rc;
    h->part_rank = rank; h->part_nranks = nranks;
    h->n_owned_entries = (int64_t)h->gpad + 4 * (int64_t)owned_halves;
    h->nvar_global = nvar_global;
    return EA_OK;
}

int ea_nccl_unique_id(const char *nccl_lib, char out[128]) {
    ea_handle *h = nullptr;
    if (!out) return EA_ERR_ARG;
    if (const char *e = load_nccl(nccl_lib)) return fail(h, EA_ERR_NCCL, "%s", e);
    static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId is expected to be 128 bytes");
    ncclUniqueId id;
    ncclResult_t r = g_nccl.GetUniqueId(&id);
    if (r != ncclSuccess) return fail(h, EA_ERR_NCCL, "ncclGetUniqueId: %s", g_nccl.GetErrorString(r));
    memcpy(out, &id, 128);
    return EA_OK;
}

int ea_comm_init(ea_handle_t *h, const char *nccl_lib, const char id[128]) {
    if (!h || !id) return EA_ERR_ARG;
    if (!h->d.partitioned) return fail(h, EA_ERR_STATE, "ea_comm_init: call ea_set_partition first");
    if (const char *e = load_nccl(nccl_lib)) return fail(h, EA_ERR_NCCL, "%s", e);
    CK(cudaSetDevice(h->device));
    ncclUniqueId uid;
    memcpy(&uid, id, 128);
    ncclResult_t r = g_nccl.CommInitRank(&h->comm, h->part_nranks, uid, h->part_rank);
    if (r != ncclSuccess) return fail(h, EA_ERR_NCCL, "ncclCommInitRank: %s", g_nccl.GetErrorString(r));
    // first collective = connection set-up (hundreds of ms): pay it here, not inside the solve
    double warm = 0.0;
    return allgather_scalar_sum(h, 0.0, &warm);
}

// Peer-memory exchange: export this rank's exchange buffer (CUDA IPC), import everybody's, switch the fused loop
// from ncclAllGather to direct NVLink stores + flags (see k_bus / k_finish). ea_comm_init is still needed for the
// handful of scalar collectives outside the inner loop.
int ea_peer_export(ea_handle_t *h, char out[64]) {
    if (!h || !out) return EA_ERR_ARG;
    if (!h->d.partitioned) return fail(h, EA_ERR_STATE, "ea_peer_export: call ea_set_partition first");
    CK(cudaSetDevice(h->device));
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "cudaIpcMemHandle_t is expected to be 64 bytes");
    if (!h->xbuf) {
        const size_t bytes = sizeof(double) * 2 * (size_t)h->part_nranks * h->d.stride + sizeof(unsigned long long) * 2 * (size_t)h->part_nranks;
        CK(cudaMalloc((void **)&h->xbuf, bytes));
        CK(cudaMemset(h->xbuf, 0, bytes));
        CK(cudaDeviceSynchronize());
    }
    cudaIpcMemHandle_t mh;
    CK(cudaIpcGetMemHandle(&mh, h->xbuf));
    memcpy(out, &mh, 64);
    return EA_OK;
}

int ea_peer_import(ea_handle_t *h, const char *handles /* nranks x 64 bytes, rank order */) {
    if (!h || !handles) return EA_ERR_ARG;
    if (!h->d.partitioned || !h->xbuf) return fail(h, EA_ERR_STATE, "ea_peer_import: call ea_peer_export first");
    CK(cudaSetDevice(h->device));
    std::vector<double *> bases((size_t)h->part_nranks, nullptr);
    for (int r = 0; r < h->part_nranks; ++r) {
        if (r == h->part_rank) { bases[r] = h->xbuf; continue; }
        cudaIpcMemHandle_t mh;
        memcpy(&mh, handles + 64 * (size_t)r, 64);
        void *p = nullptr;
        CK(cudaIpcOpenMemHandle(&p, mh, cudaIpcMemLazyEnablePeerAccess));
        h->peer_maps.push_back(p);
        bases[r] = static_cast<double *>(p);
    }
    double **dev_bases = nullptr;
    int rc = dev_upload(h, &dev_bases, bases);
    if (rc) return rc;
    h->d.peer_base = dev_bases;
    h->d.xbase = h->xbuf;
    h->d.peer_mode = 1;
    return EA_OK;
}

// Loopback exchange for single-GPU tests: the caller moves the message through the host.
// ea_part_begin = x-update + bus kernel; ea_part_get_message -> this rank's segment;
// ea_part_put_gathered <- all segments; ea_part_end = k_finish, returns the 4 norms.
int ea_part_begin(ea_handle_t *h, int64_t inner, double beta, int32_t max_auglag, double mu_max, double scale) {
    if (!h) return EA_ERR_ARG;
    if (!h->d.partitioned) return fail(h, EA_ERR_STATE, "ea_part_begin: not a partitioned handle");
    CK(cudaSetDevice(h->device));
    int rc = sync_ctrl_to_device(h, beta, -1.0, inner - 1, std::numeric_limits<long long>::max());
    if (rc) return rc;
    h->loopback = 1;
    rc = enqueue_iteration(h, max_auglag, mu_max, scale, 0);
    h->loopback = 0;
    if (rc) return rc;
    CK(cudaStreamSynchronize(h->stream));
    return EA_OK;
}
int ea_part_get_message(ea_handle_t *h, double *host, int64_t n) {
    if (!h || !host) return EA_ERR_ARG;
    if (!h->d.partitioned || n != h->d.stride) return fail(h, EA_ERR_ARG, "ea_part_get_message: n must be the stride %d", h->d.stride);
    CK(cudaSetDevice(h->device));
    CK(cudaMemcpy(host, h->d.sendbuf, sizeof(double) * (size_t)n, cudaMemcpyDeviceToHost));
    return EA_OK;
}
int ea_part_put_gathered(ea_handle_t *h, const double *host, int64_t n) {
    if (!h || !host) return EA_ERR_ARG;
    if (!h->d.partitioned || n != (int64_t)h->d.stride * h->part_nranks) return fail(h, EA_ERR_ARG, "ea_part_put_gathered: bad n");
    CK(cudaSetDevice(h->device));
    CK(cudaMemcpy(h->gather_dev, host, sizeof(double) * (size_t)n, cudaMemcpyHostToDevice));
    return EA_OK;
}
int ea_part_end(ea_handle_t *h, double out[4]) {
    if (!h || !out) return EA_ERR_ARG;
    if (!h->d.partitioned) return fail(h, EA_ERR_STATE, "ea_part_end: not a partitioned handle");
    CK(cudaSetDevice(h->device));
    k_finish<<<1, FBLOCK, 0, h->stream>>>(h->d);
    CK(cudaGetLastError());
    int rc = fetch_ctrl(h);
    if (rc) return rc;
    for (int k = 0; k < 4; ++k) out[k] = h->ctrl_host->res[k];
    return EA_OK;
}

int ea_diag_branch_eval(int device, int64_t n, const double *x, const double *param, const double *Y, double scale,
                        double *f, double *g, double *H) {
    ea_handle *h = nullptr;
    if (ea_device_count() <= 0) return fail(h, EA_ERR_CUDA, "ea_diag_branch_eval: no CUDA device");
    CK(cudaSetDevice(device));
    double *dx, *dp, *dY, *df, *dg, *dH;
    CK(cudaMalloc(&dx, 6 * n * sizeof(double))); CK(cudaMalloc(&dp, 31 * n * sizeof(double)));
    CK(cudaMalloc(&dY, 8 * n * sizeof(double))); CK(cudaMalloc(&df, n * sizeof(double)));
    CK(cudaMalloc(&dg, 6 * n * sizeof(double))); CK(cudaMalloc(&dH, 36 * n * sizeof(double)));
    CK(cudaMemcpy(dx, x, 6 * n * sizeof(double), cudaMemcpyHostToDevice));
    CK(cudaMemcpy(dp, param, 31 * n * sizeof(double), cudaMemcpyHostToDevice));
    CK(cudaMemcpy(dY, Y, 8 * n * sizeof(double), cudaMemcpyHostToDevice));
    k_diag_eval<<<nblocks(n, 128), 128>>>((int)n, dx, dp, dY, scale, df, dg, dH);
    CK(cudaGetLastError());
    CK(cudaMemcpy(f, df, n * sizeof(double), cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(g, dg, 6 * n * sizeof(double), cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(H, dH, 36 * n * sizeof(double), cudaMemcpyDeviceToHost));
    cudaFree(dx); cudaFree(dp); cudaFree(dY); cudaFree(df); cudaFree(dg); cudaFree(dH);
    return EA_OK;
}

int ea_diag_branch_solve(int device, int per_warp, int64_t n, const double *prob, int32_t max_auglag,
                         double mu_max, double scale, double *sol, int32_t *work, int64_t *cycles, double *kernel_ms) {
    ea_handle *h = nullptr;
    if (!prob || !sol || !work || !cycles || n < 0) return EA_ERR_ARG;
    if (ea_device_count() <= 0) return fail(h, EA_ERR_CUDA, "ea_diag_branch_solve: no CUDA device");
    CK(cudaSetDevice(device));
    if (n == 0) return EA_OK;
    ea_handle tmp;
    build_pow_table(&tmp, mu_max);
    double *dp, *ds; int *dw; long long *dc;
    CK(cudaMalloc(&dp, 56 * n * sizeof(double))); CK(cudaMalloc(&ds, 13 * n * sizeof(double)));
    CK(cudaMalloc(&dw, 6 * n * sizeof(int))); CK(cudaMalloc(&dc, n * sizeof(long long)));
    CK(cudaMemcpy(dp, prob, 56 * n * sizeof(double), cudaMemcpyHostToDevice));
    CK(cudaFuncSetAttribute(k_diag_solve, cudaFuncAttributeMaxDynamicSharedMemorySize, XTILE_BYTES));
    const int stride = per_warp ? 32 : 1;
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    for (int rep = 0; rep < 2; ++rep) {           // the second run is the timed one (instruction cache, clocks)
#ifdef EA_CYC
        { long long z[16] = { 0 }; CK(cudaMemcpyToSymbol(g_cyc, z, sizeof(z))); }
#endif
        CK(cudaEventRecord(e0));
        k_diag_solve<<<nblocks(n * stride, XBLOCK), XBLOCK, XTILE_BYTES>>>((int)n, stride, dp, tmp.pow_table, max_auglag, mu_max,
                                                                          scale, ds, dw, dc);
        CK(cudaEventRecord(e1));
        CK(cudaEventSynchronize(e1));
    }
    CK(cudaGetLastError());
    float ms = 0; CK(cudaEventElapsedTime(&ms, e0, e1));
    if (kernel_ms) *kernel_ms = ms;
#ifdef EA_CYC
    {   // section cycles of the long chains (>= 15 AL iterations), per AL iteration
        long long c[16]; CK(cudaMemcpyFromSymbol(c, g_cyc, sizeof(c)));
        const double al = (double)std::max(c[8], 1ll);
        fprintf(stderr, "[cyc] %lld chains, %lld AL iterations, %.2f evals/AL; per AL iteration: total %.0f | pre %.0f fg %.0f hess %.0f | "
                "newton ok %.0f (%.2f calls) newton failed %.0f literal %.0f | rest %.0f\n", c[11], c[8], c[10] / al, c[9] / al,
                c[0] / al, c[1] / al, c[2] / al, c[4] / al, c[7] / al, c[5] / al, c[6] / al,
                (c[9] - c[0] - c[1] - c[2] - c[4] - c[5] - c[6]) / al);
    }
#endif
    CK(cudaMemcpy(sol, ds, 13 * n * sizeof(double), cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(work, dw, 6 * n * sizeof(int), cudaMemcpyDeviceToHost));
    static_assert(sizeof(long long) == sizeof(int64_t), "cycle counters are 64-bit");
    CK(cudaMemcpy(cycles, dc, n * sizeof(long long), cudaMemcpyDeviceToHost));
    cudaFree(dp); cudaFree(ds); cudaFree(dw); cudaFree(dc); cudaEventDestroy(e0); cudaEventDestroy(e1);
    return EA_OK;
}

int ea_diag_fp64_peak(int device, double *tflops) {
    ea_handle *h = nullptr;
    if (!tflops) return EA_ERR_ARG;
    if (ea_device_count() <= 0) return fail(h, EA_ERR_CUDA, "ea_diag_fp64_peak: no CUDA device");
    CK(cudaSetDevice(device));
    int sms = 0;
    CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device));
    const int blocks = sms * 8, threads = 256, iters = 4096;
    double *out;
    CK(cudaMalloc(&out, sizeof(double) * blocks * threads));
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    k_fp64_peak<<<blocks, threads>>>(out, 64);           // warm-up
    double best = 0.0;
    for (int rep = 0; rep < 5; ++rep) {
        CK(cudaEventRecord(e0));
        k_fp64_peak<<<blocks, threads>>>(out, iters);
        CK(cudaEventRecord(e1));
        CK(cudaEventSynchronize(e1));
        float ms = 0; CK(cudaEventElapsedTime(&ms, e0, e1));
        const double fl = 2.0 * 64.0 * iters * (double)blocks * threads;     // 64 FMAs per loop trip
        best = std::max(best, fl / (1e-3 * ms) / 1e12);
    }
    CK(cudaGetLastError());
    cudaFree(out); cudaEventDestroy(e0); cudaEventDestroy(e1);
    *tflops = best;
    return EA_OK;
}

}  // extern "C"

#include "mp_host.inc"
#include "qp_host.inc"
#include "batch_host.inc"
