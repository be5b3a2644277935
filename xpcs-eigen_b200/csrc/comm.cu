// comm.cu -- the cross-GPU part of the path: one process (or thread) per GPU, one handle per GPU, one NCCL
// communicator rank per handle.
//
// The reference is a single-node OpenMP program; its pixel loop (corr.cpp:329-332) is what shards.  What has to
// cross GPUs:
//   1. the events.  Every GPU receives a contiguous SLAB OF FRAMES of the whole detector (its share of the file,
//      so the host->device traffic of a job is the file once, spread over all PCIe links) and owns a contiguous
//      range of pixel rows.  comm_exchange_slab() partitions the slab by owner on the device (k_demux_count /
//      k_demux_scan / k_demux_scatter: stable, frame order kept) and moves every partition to its owner with one
//      grouped ncclSend/ncclRecv over NVLink; the owner concatenates the streams of all sources in rank order =
//      frame order and ends up with an ordinary frame-major event list of its own pixels for ALL frames -- the
//      input the single-GPU ingest takes.
//   2. the per-frame sums (sparse_filter.cpp:175,190: frameSum covers all pixels), the per-static-bin sums and
//      pixelSum: element-wise SUM all-reduce after the ingest (every static bin lives on one GPU, so the sums
//      are the single-GPU values; exact for integer counts).
//   3. the normalisation partials (normalize.cu): element-wise SUM all-reduce, bit-identical to one GPU.
// NCCL is bound at run time (dlopen "libnccl.so.2"): inside a process that already carries an NCCL (PyTorch's)
// the same copy is used; the standalone host program gets the system library.  No NCCL, no multi-GPU -- the
// single-GPU path does not need it.
#include <dlfcn.h>
#include <nccl.h>
#include <unistd.h>

#include <algorithm>
#include <mutex>

#include "internal.h"

namespace xpcs {

namespace {

struct NcclApi {
    void *lib = nullptr;
    std::string error;
    decltype(&ncclGetUniqueId) GetUniqueId = nullptr;
    decltype(&ncclCommInitRank) CommInitRank = nullptr;
    decltype(&ncclCommDestroy) CommDestroy = nullptr;
    decltype(&ncclAllReduce) AllReduce = nullptr;
    decltype(&ncclAllGather) AllGather = nullptr;
    decltype(&ncclSend) Send = nullptr;
    decltype(&ncclRecv) Recv = nullptr;
    decltype(&ncclGroupStart) GroupStart = nullptr;
    decltype(&ncclGroupEnd) GroupEnd = nullptr;
    decltype(&ncclGetErrorString) GetErrorString = nullptr;
    decltype(&ncclGetVersion) GetVersion = nullptr;
};

NcclApi *nccl_api()
{
    static NcclApi api;
    static std::once_flag once;
    std::call_once(once, [] {
        // an NCCL the process already carries (PyTorch's) first; then the caller's choice; then the system library
        api.lib = dlopen("libnccl.so.2", RTLD_NOW | RTLD_LOCAL | RTLD_NOLOAD);
        const char *names[] = {getenv("XPCS_NCCL_LIB"), "libnccl.so.2", "libnccl.so"};
        for (const char *n : names) {
            if (api.lib) break;
            if (!n || !*n) continue;
            api.lib = dlopen(n, RTLD_NOW | RTLD_LOCAL);
        }
        if (!api.lib) {
            api.error = std::string("cannot load libnccl.so.2: ") + (dlerror() ? dlerror() : "not found");
            return;
        }
        bool ok = true;
        auto sym = [&](const char *name) {
            void *p = dlsym(api.lib, name);
            if (!p) {
                ok = false;
                api.error = std::string("libnccl lacks ") + name;
            }
            return p;
        };
        api.GetUniqueId = (decltype(api.GetUniqueId))sym("ncclGetUniqueId");
        api.CommInitRank = (decltype(api.CommInitRank))sym("ncclCommInitRank");
        api.CommDestroy = (decltype(api.CommDestroy))sym("ncclCommDestroy");
        api.AllReduce = (decltype(api.AllReduce))sym("ncclAllReduce");
        api.AllGather = (decltype(api.AllGather))sym("ncclAllGather");
        api.Send = (decltype(api.Send))sym("ncclSend");
        api.Recv = (decltype(api.Recv))sym("ncclRecv");
        api.GroupStart = (decltype(api.GroupStart))sym("ncclGroupStart");
        api.GroupEnd = (decltype(api.GroupEnd))sym("ncclGroupEnd");
        api.GetErrorString = (decltype(api.GetErrorString))sym("ncclGetErrorString");
        api.GetVersion = (decltype(api.GetVersion))sym("ncclGetVersion");
        if (!ok) {
            dlclose(api.lib);
            api.lib = nullptr;
        }
    });
    return api.lib ? &api : nullptr;
}

int nccl_check(xpcs_handle_s *h, ncclResult_t r, const char *what)
{
    if (r == ncclSuccess) return XPCS_OK;
    NcclApi *n = nccl_api();
    return fail(h, XPCS_E_CUDA, "%s: NCCL error %d (%s)", what, (int)r, n ? n->GetErrorString(r) : "?");
}

constexpr int kMaxRanks = 32;
constexpr int kDmWarps = 8;

// Pass 1 of the partition: one warp per frame of the slab counts the frame's events per owner.
// cnt[d * (nfr + 1) + j] = events of slab frame j that belong to shard d (masked pixels belong to nobody).
__global__ void __launch_bounds__(kDmWarps * 32) k_demux_count(const int32_t *__restrict__ idx, const int64_t *__restrict__ off,
                                                               int nfr, const unsigned char *__restrict__ owner, int P, int N,
                                                               int64_t *__restrict__ cnt)
{
    __shared__ int sc[kDmWarps][32];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int j = blockIdx.x * kDmWarps + warp;
    if (j >= nfr) return;  // warp-uniform; no block barrier below
    sc[warp][lane] = 0;
    __syncwarp();
    const int64_t a = off[j], b = off[j + 1];
    for (int64_t e0 = a; e0 < b; e0 += 32) {
        const int64_t e = e0 + lane;
        int o = 255;
        if (e < b) {
            const int pix = idx[e];
            if ((unsigned)pix < (unsigned)P) o = owner[pix];
        }
        const unsigned peers = __match_any_sync(0xffffffffu, o);
        if (o < N && lane == __ffs(peers) - 1) sc[warp][o] += __popc(peers);
        __syncwarp();
    }
    if (lane < N) cnt[(int64_t)lane * (nfr + 1) + j] = (int64_t)sc[warp][lane];
}

// Exclusive scan over the slab's frames, one CTA per destination, in place; entry nfr = the total.
// meta = [first frame of the slab, frames of the slab, events for shard 0, 1, ...]: what every rank tells the others.
__global__ void __launch_bounds__(1024) k_demux_scan(int64_t *__restrict__ cnt, int nfr, int first, int64_t *__restrict__ meta)
{
    __shared__ long long part[1024];
    const int d = blockIdx.x, tid = threadIdx.x;
    int64_t *c = cnt + (int64_t)d * (nfr + 1);
    const int per = (nfr + 1023) / 1024;
    const int a = min(nfr, tid * per), b = min(nfr, a + per);
    long long s = 0;
    for (int i = a; i < b; i++) s += c[i];
    part[tid] = s;
    __syncthreads();
    for (int o = 1; o < 1024; o <<= 1) {
        const long long x = tid >= o ? part[tid - o] : 0;
        __syncthreads();
        part[tid] += x;
        __syncthreads();
    }
    long long run = part[tid] - s;
    for (int i = a; i < b; i++) {
        const long long v = c[i];
        c[i] = run;
        run += v;
    }
    if (tid == 1023) {
        c[nfr] = part[1023];
        meta[2 + d] = part[1023];
        if (d == 0) {
            meta[0] = first;
            meta[1] = nfr;
        }
    }
}

struct DemuxDst {
    int32_t *idx[kMaxRanks];  // where the stream for shard d starts (send buffer, or the final list for the own shard)
    int16_t *val[kMaxRanks];
};

// Pass 2: the same walk, every event goes to position (frame offset of its owner's stream + rank inside the
// frame); stable, so a stream is the slab restricted to the owner's pixels, frame by frame, pixel order kept.
__global__ void __launch_bounds__(kDmWarps * 32) k_demux_scatter(const int32_t *__restrict__ idx, const int16_t *__restrict__ val,
                                                                 const int64_t *__restrict__ off, int nfr,
                                                                 const unsigned char *__restrict__ owner, int P, int N,
                                                                 const int64_t *__restrict__ dm_off, DemuxDst dst)
{
    __shared__ long long cur[kDmWarps][32];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int j = blockIdx.x * kDmWarps + warp;
    if (j >= nfr) return;
    cur[warp][lane] = lane < N ? dm_off[(int64_t)lane * (nfr + 1) + j] : 0;
    __syncwarp();
    const int64_t a = off[j], b = off[j + 1];
    for (int64_t e0 = a; e0 < b; e0 += 32) {
        const int64_t e = e0 + lane;
        int o = 255, pix = 0;
        int16_t v = 0;
        if (e < b) {
            pix = idx[e];
            v = val[e];
            if ((unsigned)pix < (unsigned)P) o = owner[pix];
        }
        const unsigned peers = __match_any_sync(0xffffffffu, o);
        const int before = __popc(peers & ((1u << lane) - 1u));
        if (o < N) {
            const long long pos = cur[warp][o] + before;
            dst.idx[o][pos] = pix;
            dst.val[o][pos] = v;
        }
        __syncwarp();
        if (o < N && before == 0) cur[warp][o] += __popc(peers);
        __syncwarp();
    }
}

// Direct transport: the same partition, but a warp first gathers up to kP2pChunk events of its frame by owner in
// its own shared-memory area and then copies every owner's run out with consecutive lanes on consecutive
// addresses.  Storing event by event (k_demux_scatter on peer pointers) puts ~4 events = 16 + 8 bytes into every
// NVLink write and stalls at the link's packet rate (0.64 ms for a 79 MB slab at 8 GPUs); runs of ~128 events per
// owner travel as 128-byte writes.  Order inside an owner's stream is unchanged (stable), so the result is
// bit-identical to the staged path.
constexpr int kP2pChunk = 1024;
constexpr int kRunWarps = 4;  // 4 x 6 KB of chunk areas per CTA

__global__ void __launch_bounds__(kRunWarps * 32) k_demux_scatter_runs(const int32_t *__restrict__ idx, const int16_t *__restrict__ val,
                                                                      const int64_t *__restrict__ off, int nfr,
                                                                      const unsigned char *__restrict__ owner, int P, int N,
                                                                      const int64_t *__restrict__ dm_off, DemuxDst dst)
{
    __shared__ int32_t s_idx[kRunWarps][kP2pChunk];
    __shared__ int16_t s_val[kRunWarps][kP2pChunk];
    __shared__ long long cur[kRunWarps][32];  // position of the frame's next event in every owner's stream
    __shared__ int cnt[kRunWarps][32];        // events of the chunk per owner
    __shared__ int sb[kRunWarps][32];         // where an owner's run starts in the chunk area
    __shared__ int fill[kRunWarps][32];       // events of the owner placed so far
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int j = blockIdx.x * kRunWarps + warp;
    if (j >= nfr) return;  // warp-uniform; no block barrier below
    cur[warp][lane] = lane < N ? dm_off[(int64_t)lane * (nfr + 1) + j] : 0;
    const int64_t a = off[j], b = off[j + 1];
    for (int64_t c0 = a; c0 < b; c0 += kP2pChunk) {
        const int64_t c1 = c0 + kP2pChunk < b ? c0 + kP2pChunk : b;
        cnt[warp][lane] = 0;
        fill[warp][lane] = 0;
        __syncwarp();
        // pass A: events of the chunk per owner
        for (int64_t e0 = c0; e0 < c1; e0 += 32) {
            const int64_t e = e0 + lane;
            int o = 255;
            if (e < c1) {
                const int pix = idx[e];
                if ((unsigned)pix < (unsigned)P) o = owner[pix];
            }
            const unsigned peers = __match_any_sync(0xffffffffu, o);
            if (o < N && lane == __ffs(peers) - 1) cnt[warp][o] += __popc(peers);
            __syncwarp();
        }
        {   // exclusive prefix over the owners (lane = owner)
            const int c = lane < N ? cnt[warp][lane] : 0;
            int x = c;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int y = __shfl_up_sync(0xffffffffu, x, o);
                if (lane >= o) x += y;
            }
            sb[warp][lane] = x - c;
        }
        __syncwarp();
        // pass B: gather by owner, order kept
        for (int64_t e0 = c0; e0 < c1; e0 += 32) {
            const int64_t e = e0 + lane;
            int o = 255, pix = 0;
            int16_t v = 0;
            if (e < c1) {
                pix = idx[e];
                v = val[e];
                if ((unsigned)pix < (unsigned)P) o = owner[pix];
            }
            const unsigned peers = __match_any_sync(0xffffffffu, o);
            const int before = __popc(peers & ((1u << lane) - 1u));
            if (o < N) {
                const int p = sb[warp][o] + fill[warp][o] + before;
                s_idx[warp][p] = pix;
                s_val[warp][p] = v;
            }
            __syncwarp();
            if (o < N && before == 0) fill[warp][o] += __popc(peers);
            __syncwarp();
        }
        // pass C: every owner's run leaves with consecutive lanes on consecutive addresses
        for (int d = 0; d < N; d++) {
            const int n = cnt[warp][d], s0 = sb[warp][d];
            const long long g0 = cur[warp][d];
            int32_t *gi = dst.idx[d] + g0;
            int16_t *gv = dst.val[d] + g0;
            for (int k = lane; k < n; k += 32) {
                gi[k] = s_idx[warp][s0 + k];
                gv[k] = s_val[warp][s0 + k];
            }
        }
        __syncwarp();
        if (lane < N) cur[warp][lane] += cnt[warp][lane];
        __syncwarp();
    }
}

struct OffDst {
    int64_t *p[kMaxRanks];
};

// direct path: the per-destination frame offsets of this slab go straight into the owners' offset tables
__global__ void k_push_offsets(const int64_t *__restrict__ dm_off, int nfr, OffDst dst)
{
    const int64_t *src = dm_off + (int64_t)blockIdx.x * (nfr + 1);
    int64_t *d = dst.p[blockIdx.x];
    for (int i = blockIdx.y * blockDim.x + threadIdx.x; i <= nfr; i += gridDim.y * blockDim.x) d[i] = src[i];
}

// what a rank tells the others about its receive buffers (28 x int64)
struct PeerRec {
    unsigned long long idx, val, off, pid;
    cudaIpcMemHandle_t hidx, hval, hoff;
};
static_assert(sizeof(PeerRec) == 224, "PeerRec is 28 int64");

struct MergeRanges {
    int first[kMaxRanks], nfr[kMaxRanks];
    long long base[kMaxRanks];
};

// received per-source offsets -> one frame-offset array over all raw frames of the job
__global__ void k_merge_offsets(const int64_t *__restrict__ recv_off, int64_t *__restrict__ frame_off, MergeRanges r, int N,
                                int raw_total, long long total)
{
    const int f = blockIdx.x * blockDim.x + threadIdx.x;
    if (f > raw_total) return;
    if (f == raw_total) {
        frame_off[f] = total;
        return;
    }
    int s = 0;
    while (s + 1 < N && f >= r.first[s] + r.nfr[s]) s++;
    frame_off[f] = r.base[s] + recv_off[(int64_t)r.first[s] + s + (f - r.first[s])];
}

}  // namespace

bool comm_active(const xpcs_handle_s *h) { return h->comm != nullptr && h->comm_nranks > 1; }

int comm_allreduce_f64(xpcs_handle_s *h, double *d_buf, size_t n)
{
    if (!comm_active(h) || n == 0) return XPCS_OK;
    NcclApi *api = nccl_api();
    LaunchScope ls(h, "nccl_allreduce", false);
    return nccl_check(h, api->AllReduce(d_buf, d_buf, n, ncclFloat64, ncclSum, (ncclComm_t)h->comm, h->stream), "ncclAllReduce");
}

int comm_allreduce_f32(xpcs_handle_s *h, float *d_buf, size_t n)
{
    if (!comm_active(h) || n == 0) return XPCS_OK;
    NcclApi *api = nccl_api();
    LaunchScope ls(h, "nccl_allreduce", false);
    return nccl_check(h, api->AllReduce(d_buf, d_buf, n, ncclFloat32, ncclSum, (ncclComm_t)h->comm, h->stream), "ncclAllReduce");
}

// several in-place SUM all-reduces of doubles as ONE grouped NCCL launch (small buffers: launch latency dominates)
int comm_allreduce_f64_group(xpcs_handle_s *h, double *const *bufs, const size_t *counts, int n)
{
    if (!comm_active(h)) return XPCS_OK;
    NcclApi *api = nccl_api();
    LaunchScope ls(h, "nccl_allreduce", false);
    ncclResult_t r = api->GroupStart();
    for (int i = 0; i < n && r == ncclSuccess; i++)
        if (counts[i] > 0) r = api->AllReduce(bufs[i], bufs[i], counts[i], ncclFloat64, ncclSum, (ncclComm_t)h->comm, h->stream);
    const ncclResult_t r2 = api->GroupEnd();
    return nccl_check(h, r != ncclSuccess ? r : r2, "grouped ncclAllReduce");
}

void comm_destroy(xpcs_handle_s *h)
{
    if (h->comm) {
        NcclApi *api = nccl_api();
        if (api) api->CommDestroy((ncclComm_t)h->comm);
        h->comm = nullptr;
    }
    for (void *p : h->peer_opened) cudaIpcCloseMemHandle(p);
    h->peer_opened.clear();
    release(h->d_p2p_xchg);
    release(h->d_owner_of_pixel);
    release(h->d_slab_idx); release(h->d_slab_val); release(h->d_slab_off);
    release(h->d_dm_off); release(h->d_dm_meta); release(h->d_dm_all);
    release(h->d_send_idx); release(h->d_send_val); release(h->d_recv_off);
}

// Make the receive buffers of every rank addressable from this device: same process -> the pointers themselves
// (peer access enabled at xpcs_comm_init), other processes -> CUDA IPC mappings.  Collective.  On any failure
// anywhere, every rank drops to the staged ncclSend/ncclRecv path for good (the decision is all-reduced).
static int p2p_exchange_mappings(xpcs_handle_s *h)
{
    const int N = h->comm_nranks, me = h->comm_rank;
    NcclApi *api = nccl_api();
    int rc;
    constexpr size_t kRecWords = sizeof(PeerRec) / 8;
    if ((rc = ensure(h, h->d_p2p_xchg, (size_t)(N + 1) * kRecWords + 8, "peer mapping records"))) return rc;
    PeerRec mine;
    memset(&mine, 0, sizeof(mine));
    mine.idx = (unsigned long long)(uintptr_t)h->d_idx.p;
    mine.val = (unsigned long long)(uintptr_t)h->d_val.p;
    mine.off = (unsigned long long)(uintptr_t)h->d_recv_off.p;
    mine.pid = (unsigned long long)getpid();
    // IPC handles only when some rank lives in another process (threads of one process share the address space)
    bool need_ipc = false;
    for (int s = 0; s < N; s++) need_ipc = need_ipc || ((int)h->peer_pid.size() == N && h->peer_pid[s] != (long long)mine.pid);
    bool ok = true;
    if (need_ipc) {
        ok = cudaIpcGetMemHandle(&mine.hidx, h->d_idx.p) == cudaSuccess && cudaIpcGetMemHandle(&mine.hval, h->d_val.p) == cudaSuccess &&
             cudaIpcGetMemHandle(&mine.hoff, h->d_recv_off.p) == cudaSuccess;
        if (!ok) cudaGetLastError();
    }
    int64_t *d_mine = h->d_p2p_xchg.p, *d_all = h->d_p2p_xchg.p + kRecWords;
    if ((rc = check_cuda(h, cudaMemcpyAsync(d_mine, &mine, sizeof(mine), cudaMemcpyHostToDevice, h->stream), "peer record H2D"))) return rc;
    {
        LaunchScope ls(h, "nccl_allgather", false);
        if ((rc = nccl_check(h, api->AllGather(d_mine, d_all, kRecWords, ncclInt64, (ncclComm_t)h->comm, h->stream), "ncclAllGather (peer records)")))
            return rc;
    }
    std::vector<PeerRec> all((size_t)N);
    if ((rc = check_cuda(h, cudaMemcpyAsync(all.data(), d_all, sizeof(PeerRec) * (size_t)N, cudaMemcpyDeviceToHost, h->stream), "peer records D2H")))
        return rc;
    if ((rc = check_cuda(h, cudaStreamSynchronize(h->stream), "peer records"))) return rc;
    for (void *p : h->peer_opened) cudaIpcCloseMemHandle(p);
    h->peer_opened.clear();
    h->peer_idx.assign(N, nullptr);
    h->peer_val.assign(N, nullptr);
    h->peer_off.assign(N, nullptr);
    for (int s = 0; s < N && ok; s++) {
        if (s == me) {
            h->peer_idx[s] = h->d_idx.p;
            h->peer_val[s] = h->d_val.p;
            h->peer_off[s] = h->d_recv_off.p;
        } else if (all[s].pid == mine.pid) {  // another thread of this process: one address space
            h->peer_idx[s] = reinterpret_cast<int32_t *>((uintptr_t)all[s].idx);
            h->peer_val[s] = reinterpret_cast<int16_t *>((uintptr_t)all[s].val);
            h->peer_off[s] = reinterpret_cast<int64_t *>((uintptr_t)all[s].off);
        } else {
            void *pi = nullptr, *pv = nullptr, *po = nullptr;
            ok = cudaIpcOpenMemHandle(&pi, all[s].hidx, cudaIpcMemLazyEnablePeerAccess) == cudaSuccess;
            if (ok) h->peer_opened.push_back(pi);
            ok = ok && cudaIpcOpenMemHandle(&pv, all[s].hval, cudaIpcMemLazyEnablePeerAccess) == cudaSuccess;
            if (ok) h->peer_opened.push_back(pv);
            ok = ok && cudaIpcOpenMemHandle(&po, all[s].hoff, cudaIpcMemLazyEnablePeerAccess) == cudaSuccess;
            if (ok) h->peer_opened.push_back(po);
            if (!ok) cudaGetLastError();
            h->peer_idx[s] = (int32_t *)pi;
            h->peer_val[s] = (int16_t *)pv;
            h->peer_off[s] = (int64_t *)po;
        }
    }
    // all or nothing: MIN over the ranks of "every mapping worked here"
    int64_t flag = ok ? 1 : 0;
    if ((rc = check_cuda(h, cudaMemcpyAsync(d_mine, &flag, sizeof(flag), cudaMemcpyHostToDevice, h->stream), "p2p flag H2D"))) return rc;
    {
        LaunchScope ls(h, "nccl_allreduce", false);
        if ((rc = nccl_check(h, api->AllReduce(d_mine, d_mine, 1, ncclInt64, ncclMin, (ncclComm_t)h->comm, h->stream), "ncclAllReduce (p2p)")))
            return rc;
    }
    if ((rc = check_cuda(h, cudaMemcpyAsync(&flag, d_mine, sizeof(flag), cudaMemcpyDeviceToHost, h->stream), "p2p flag D2H"))) return rc;
    if ((rc = check_cuda(h, cudaStreamSynchronize(h->stream), "p2p agreement"))) return rc;
    if (!flag) {
        for (void *p : h->peer_opened) cudaIpcCloseMemHandle(p);
        h->peer_opened.clear();
        h->p2p_enabled = false;
        h->p2p_mapped = false;
        return XPCS_OK;
    }
    h->p2p_mapped = true;
    return XPCS_OK;
}

// Frame slabs -> pixel shards.  On return the handle holds a frame-major event list of its own pixels over all
// raw frames of the job (ev_idx / ev_val / ev_off), exactly what xpcs_push_sparse_device would have set up.
// Two transports: DIRECT (default on one host with peer access) -- the partition kernel stores every event
// straight into its owner's list over NVLink, the transfer IS the partition; STAGED -- per-owner streams in a send
// buffer, one grouped ncclSend/ncclRecv (XPCS_NO_P2P, or whenever peer mappings cannot be had).
int comm_exchange_slab(xpcs_handle_s *h)
{
    const int N = h->comm_nranks, me = h->comm_rank, nfr = h->slab_frames;
    const int W = N + 6;  // table row: first frame, frames, events per destination [N], receive capacities (events, offsets), buffer generation, wants mappings
    NcclApi *api = nccl_api();
    int rc;
    if ((rc = ensure(h, h->d_dm_off, (size_t)N * (nfr + 1), "slab partition offsets"))) return rc;
    if ((rc = ensure(h, h->d_dm_meta, (size_t)W, "slab partition totals"))) return rc;
    if ((rc = ensure(h, h->d_dm_all, (size_t)N * W, "slab partition table"))) return rc;
    cudaMemsetAsync(h->d_dm_off.p, 0, sizeof(int64_t) * (size_t)N * (nfr + 1), h->stream);
    if (nfr > 0) {
        LaunchScope ls(h, "k_demux_count");
        k_demux_count<<<(nfr + kDmWarps - 1) / kDmWarps, kDmWarps * 32, 0, h->stream>>>(
            h->slab_idx, h->slab_off, nfr, h->d_owner_of_pixel.p, h->P, N, h->d_dm_off.p);
    }
    {
        LaunchScope ls(h, "k_demux_scan");
        k_demux_scan<<<N, 1024, 0, h->stream>>>(h->d_dm_off.p, nfr, h->slab_first, h->d_dm_meta.p);
    }
    {
        // buffers that moved outside an exchange (a plain push grew them) invalidate the peers' mappings as well
        if (h->p2p_mapped && (int)h->peer_idx.size() == N &&
            (h->peer_idx[me] != h->d_idx.p || h->peer_val[me] != h->d_val.p || h->peer_off[me] != h->d_recv_off.p))
            h->p2p_gen++;
        const int64_t extra[4] = {h->d_idx.p && h->d_val.p ? (int64_t)std::min(h->d_idx.n, h->d_val.n) : 0,
                                  h->d_recv_off.p ? (int64_t)h->d_recv_off.n : 0, (int64_t)h->p2p_gen, h->p2p_mapped ? 0 : 1};
        if ((rc = check_cuda(h, cudaMemcpyAsync(h->d_dm_meta.p + N + 2, extra, sizeof(extra), cudaMemcpyHostToDevice, h->stream), "table H2D")))
            return rc;
        LaunchScope ls(h, "nccl_allgather", false);
        if ((rc = nccl_check(h, api->AllGather(h->d_dm_meta.p, h->d_dm_all.p, (size_t)W, ncclInt64, (ncclComm_t)h->comm, h->stream),
                             "ncclAllGather")))
            return rc;
    }
    std::vector<int64_t> all((size_t)N * W);
    if ((rc = check_cuda(h, cudaMemcpyAsync(all.data(), h->d_dm_all.p, sizeof(int64_t) * all.size(), cudaMemcpyDeviceToHost, h->stream),
                         "slab table D2H")))
        return rc;
    if ((rc = check_cuda(h, cudaStreamSynchronize(h->stream), "slab partition"))) return rc;
    auto at = [&](int s, int k) { return all[(size_t)s * W + k]; };
    // the slabs must tile the raw frames of the job in rank order
    MergeRanges mr{};
    int64_t raw_total = 0, E_me = 0;
    h->slab_first_of_rank.assign(N, 0);
    h->slab_frames_of_rank.assign(N, 0);
    for (int s = 0; s < N; s++) {
        if (at(s, 0) != raw_total)
            return fail(h, XPCS_E_ARG, "frame slabs do not tile the job: rank %d starts at frame %lld, expected %lld", s,
                        (long long)at(s, 0), (long long)raw_total);
        mr.first[s] = (int)at(s, 0);
        mr.nfr[s] = (int)at(s, 1);
        mr.base[s] = E_me;
        h->slab_first_of_rank[s] = mr.first[s];
        h->slab_frames_of_rank[s] = mr.nfr[s];
        raw_total += at(s, 1);
        E_me += at(s, 2 + me);
    }
    if (raw_total > 0x7fffffffLL) return fail(h, XPCS_E_ARG, "too many raw frames");
    // The receive buffers grow only when the capacities published in the table say so -- every rank can then tell
    // from the table alone whose buffers move in this exchange (the peers' mappings must follow).
    if (at(me, N + 2) < E_me + 8) {
        size_t want = (size_t)E_me + (size_t)E_me / 16;
        if (h->prm.reserve_events > (int64_t)want) want = (size_t)h->prm.reserve_events;
        release(h->d_idx);
        release(h->d_val);
        if ((rc = ensure(h, h->d_idx, want + 8, "event indices"))) return rc;
        if ((rc = ensure(h, h->d_val, want + 8, "event values"))) return rc;
        h->p2p_gen++;
    }
    if (at(me, N + 3) < raw_total + N + 1) {
        release(h->d_recv_off);
        if ((rc = ensure(h, h->d_recv_off, (size_t)raw_total + N + 1, "received offsets"))) return rc;
        if (!(at(me, N + 2) < E_me + 8)) h->p2p_gen++;  // (one bump per exchange, whatever moved)
    }
    if ((rc = ensure(h, h->d_frame_off, (size_t)raw_total + 1, "frame offsets"))) return rc;

    bool direct = h->p2p_enabled;
    if (direct) {
        // Every rank derives from the table whether ANY receive buffer moves in this exchange (the same answer
        // everywhere): then, or when somebody still lacks mappings, the mapping records go round once more.
        bool remap = false;
        for (int s = 0; s < N; s++) {
            int64_t E_s = 0;
            for (int q = 0; q < N; q++) E_s += at(q, 2 + s);
            const bool grows = at(s, N + 2) < E_s + 8 || at(s, N + 3) < raw_total + N + 1;
            const bool unseen = (int)h->peer_gen_seen.size() != N || h->peer_gen_seen[s] != at(s, N + 4);
            remap = remap || grows || unseen || at(s, N + 5) != 0;
        }
        if (remap) {
            if ((rc = p2p_exchange_mappings(h))) return rc;
            h->peer_gen_seen.assign(N, -1);
            // generations after this exchange: a rank that had to grow has bumped its own by exactly one
            for (int s = 0; s < N; s++) {
                int64_t E_s = 0;
                for (int q = 0; q < N; q++) E_s += at(q, 2 + s);
                const bool grows = at(s, N + 2) < E_s + 8 || at(s, N + 3) < raw_total + N + 1;
                h->peer_gen_seen[s] = at(s, N + 4) + (grows ? 1 : 0);
            }
            h->peer_gen_seen[me] = h->p2p_gen;
        }
        direct = h->p2p_enabled && h->p2p_mapped;
    }

    DemuxDst dst{};
    int64_t send_total = 0;
    std::vector<int64_t> sbase(N, 0);
    if (direct) {
        // position of my stream inside owner d's list: behind the streams of the ranks before me
        for (int d = 0; d < N; d++) {
            int64_t base = 0;
            for (int q = 0; q < me; q++) base += at(q, 2 + d);
            dst.idx[d] = h->peer_idx[d] + base;
            dst.val[d] = h->peer_val[d] + base;
        }
    } else {
        for (int d = 0; d < N; d++) {
            sbase[d] = send_total;
            if (d != me) send_total += at(me, 2 + d);
        }
        // the send streams live in the record buffer of the store build, which is only written after the exchange
        // (6 bytes per event to send against 8 bytes per event to keep)
        const size_t send_words = ((size_t)send_total * 6 + 64) / 8 + 8;
        if ((rc = ensure(h, h->d_rec, std::max(send_words, (size_t)E_me + 1), "event records / send streams"))) return rc;
        int32_t *send_idx = reinterpret_cast<int32_t *>(h->d_rec.p);
        int16_t *send_val = reinterpret_cast<int16_t *>(send_idx + (((size_t)send_total + 7) & ~(size_t)7));
        for (int d = 0; d < N; d++) {
            if (d == me) {
                dst.idx[d] = h->d_idx.p + mr.base[me];
                dst.val[d] = h->d_val.p + mr.base[me];
            } else {
                dst.idx[d] = send_idx + sbase[d];
                dst.val[d] = send_val + sbase[d];
            }
        }
    }
    if (nfr > 0) {
        LaunchScope ls(h, direct ? "k_demux_scatter_p2p" : "k_demux_scatter");
        // event-by-event stores carry 32 / N events per owner and warp store: wide enough for two ranks (measured 0.57
        // against 0.91 ms with the run gather, whose three passes over the slab then dominate), too narrow from four on
        // (0.68 / 0.64 against 0.53 / 0.35 ms at 4 / 8 GPUs)
        // (XPCS_DEMUX_RUNS = 1 / 0 forces one kernel or the other, also on the staged path: the tests run both on one GPU)
        bool runs = direct && N >= 4;
        if (const char *e = getenv("XPCS_DEMUX_RUNS")) runs = atoi(e) != 0;
        if (runs)
            k_demux_scatter_runs<<<(nfr + kRunWarps - 1) / kRunWarps, kRunWarps * 32, 0, h->stream>>>(
                h->slab_idx, h->slab_val, h->slab_off, nfr, h->d_owner_of_pixel.p, h->P, N, h->d_dm_off.p, dst);
        else
            k_demux_scatter<<<(nfr + kDmWarps - 1) / kDmWarps, kDmWarps * 32, 0, h->stream>>>(
                h->slab_idx, h->slab_val, h->slab_off, nfr, h->d_owner_of_pixel.p, h->P, N, h->d_dm_off.p, dst);
    }
    if (direct) {
        OffDst od{};
        for (int d = 0; d < N; d++) od.p[d] = h->peer_off[d] + mr.first[me] + me;
        {
            LaunchScope ls(h, "k_push_offsets");
            k_push_offsets<<<dim3(N, 16), 256, 0, h->stream>>>(h->d_dm_off.p, nfr, od);
        }
        // nobody reads its list before every rank has finished storing into it: a one-word all-reduce as the barrier
        LaunchScope ls(h, "nccl_barrier", false);
        if ((rc = nccl_check(h, api->AllReduce(h->d_dm_all.p, h->d_dm_all.p, 1, ncclInt64, ncclSum, (ncclComm_t)h->comm, h->stream),
                             "ncclAllReduce (barrier)")))
            return rc;
    } else {
        cudaMemcpyAsync(h->d_recv_off.p + mr.first[me] + me, h->d_dm_off.p + (size_t)me * (nfr + 1), sizeof(int64_t) * ((size_t)nfr + 1),
                        cudaMemcpyDeviceToDevice, h->stream);
        LaunchScope ls(h, "nccl_exchange", false);
        ncclComm_t comm = (ncclComm_t)h->comm;
        ncclResult_t r = api->GroupStart();
        for (int s = 0; s < N && r == ncclSuccess; s++) {
            if (s == me) continue;
            const int64_t ns = at(me, 2 + s), nr = at(s, 2 + me);
            if (ns > 0) {
                r = api->Send(dst.idx[s], (size_t)ns, ncclInt32, s, comm, h->stream);
                if (r == ncclSuccess) r = api->Send(dst.val[s], (size_t)ns * 2, ncclInt8, s, comm, h->stream);
            }
            if (r == ncclSuccess)
                r = api->Send(h->d_dm_off.p + (size_t)s * (nfr + 1), (size_t)nfr + 1, ncclInt64, s, comm, h->stream);
            if (nr > 0 && r == ncclSuccess) {
                r = api->Recv(h->d_idx.p + mr.base[s], (size_t)nr, ncclInt32, s, comm, h->stream);
                if (r == ncclSuccess) r = api->Recv(h->d_val.p + mr.base[s], (size_t)nr * 2, ncclInt8, s, comm, h->stream);
            }
            if (r == ncclSuccess)
                r = api->Recv(h->d_recv_off.p + mr.first[s] + s, (size_t)mr.nfr[s] + 1, ncclInt64, s, comm, h->stream);
        }
        const ncclResult_t r2 = api->GroupEnd();
        if ((rc = nccl_check(h, r != ncclSuccess ? r : r2, "event exchange (ncclSend/ncclRecv)"))) return rc;
    }
    {
        LaunchScope ls(h, "k_merge_offsets");
        k_merge_offsets<<<(int)((raw_total + 1 + 255) / 256), 256, 0, h->stream>>>(h->d_recv_off.p, h->d_frame_off.p, mr, N,
                                                                                 (int)raw_total, (long long)E_me);
    }
    h->ev_idx = h->d_idx.p;
    h->ev_val = h->d_val.p;
    h->ev_off = h->d_frame_off.p;
    h->E = E_me;
    h->raw_frames = (int)raw_total;
    h->external_events = true;
    h->ts_clock.resize((size_t)raw_total, 0.0);
    h->ts_ticks.resize((size_t)raw_total, 0.0);
    return check_cuda(h, cudaGetLastError(), "slab exchange kernels");
}

}  // namespace xpcs

using namespace xpcs;

extern "C" int xpcs_comm_unique_id(void *id128)
{
    if (!id128) return XPCS_E_ARG;
    NcclApi *api = nccl_api();
    if (!api) return fail(nullptr, XPCS_E_CUDA, "NCCL unavailable");
    static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId is 128 bytes");
    ncclUniqueId id;
    ncclResult_t r = api->GetUniqueId(&id);
    if (r != ncclSuccess) return fail(nullptr, XPCS_E_CUDA, "ncclGetUniqueId: %s", api->GetErrorString(r));
    memcpy(id128, &id, sizeof(id));
    return XPCS_OK;
}

extern "C" int xpcs_comm_init(xpcs_handle h, int nranks, int rank, const void *id128)
{
    if (!h || !id128) return h ? fail(h, XPCS_E_ARG, "comm_init: bad arguments") : XPCS_E_ARG;
    if (nranks != h->prm.shard_count || rank != h->prm.shard_index)
        return fail(h, XPCS_E_ARG, "comm_init: rank %d of %d does not match the handle's shard %d of %d", rank, nranks,
                    h->prm.shard_index, h->prm.shard_count);
    if (nranks > kMaxRanks) return fail(h, XPCS_E_ARG, "comm_init: at most %d ranks", kMaxRanks);
    if (h->comm) return fail(h, XPCS_E_STATE, "comm_init called twice");
    NcclApi *api = nccl_api();
    if (!api) return fail(h, XPCS_E_CUDA, "NCCL unavailable (libnccl.so.2 not loadable)");
    cudaSetDevice(h->device);
    ncclUniqueId id;
    memcpy(&id, id128, sizeof(id));
    ncclComm_t comm = nullptr;
    int rc = nccl_check(h, api->CommInitRank(&comm, nranks, id, rank), "ncclCommInitRank");
    if (rc) return rc;
    h->comm = comm;
    h->comm_nranks = nranks;
    h->comm_rank = rank;
    if ((rc = ensure(h, h->d_owner_of_pixel, (size_t)h->P, "pixel owners"))) return rc;
    if ((rc = check_cuda(h, cudaMemcpy(h->d_owner_of_pixel.p, h->owner_of_pixel.data(), (size_t)h->P, cudaMemcpyHostToDevice),
                         "pixel owners H2D")))
        return rc;
    // Direct NVLink stores need every rank on this host and peer access between the devices.  Who is where is
    // all-gathered once; the verdict is all-reduced so that every rank takes the same transport.
    h->p2p_enabled = false;
    h->p2p_mapped = false;
    if (nranks > 1) {
        int64_t mine[3] = {(int64_t)getpid(), (int64_t)gethostid(), (int64_t)h->device};
        if ((rc = ensure(h, h->d_p2p_xchg, (size_t)3 * (nranks + 1) + 8, "peer identities"))) return rc;
        int64_t *d_mine = h->d_p2p_xchg.p, *d_all = h->d_p2p_xchg.p + 3;
        std::vector<int64_t> all((size_t)3 * nranks);
        cudaMemcpyAsync(d_mine, mine, sizeof(mine), cudaMemcpyHostToDevice, h->stream);
        if ((rc = nccl_check(h, api->AllGather(d_mine, d_all, 3, ncclInt64, comm, h->stream), "ncclAllGather (identities)"))) return rc;
        cudaMemcpyAsync(all.data(), d_all, sizeof(int64_t) * all.size(), cudaMemcpyDeviceToHost, h->stream);
        if ((rc = check_cuda(h, cudaStreamSynchronize(h->stream), "peer identities"))) return rc;
        bool ok = !getenv("XPCS_NO_P2P");
        h->peer_pid.assign(nranks, 0);
        for (int s = 0; s < nranks && ok; s++) {
            h->peer_pid[s] = all[(size_t)3 * s];
            if (all[(size_t)3 * s + 1] != mine[1]) ok = false;  // another host: NCCL only
            if (s != rank && all[(size_t)3 * s] == mine[0]) {   // another thread of this process
                const int dev = (int)all[(size_t)3 * s + 2];
                int can = 0;
                if (dev == h->device || cudaDeviceCanAccessPeer(&can, h->device, dev) != cudaSuccess || !can) ok = false;
                else {
                    cudaError_t e = cudaDeviceEnablePeerAccess(dev, 0);
                    if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) ok = false;
                    cudaGetLastError();
                }
            }
        }
        int64_t flag = ok ? 1 : 0;
        cudaMemcpyAsync(d_mine, &flag, sizeof(flag), cudaMemcpyHostToDevice, h->stream);
        if ((rc = nccl_check(h, api->AllReduce(d_mine, d_mine, 1, ncclInt64, ncclMin, comm, h->stream), "ncclAllReduce (transport)"))) return rc;
        cudaMemcpyAsync(&flag, d_mine, sizeof(flag), cudaMemcpyDeviceToHost, h->stream);
        if ((rc = check_cuda(h, cudaStreamSynchronize(h->stream), "transport agreement"))) return rc;
        h->p2p_enabled = flag != 0;
    }
    return XPCS_OK;
}

/* 1 = the slab exchange of this handle stores directly into the owners' buffers over NVLink (peer mappings),
 * 0 = it stages and uses ncclSend/ncclRecv, -1 = no communicator */
extern "C" int xpcs_comm_transport(xpcs_handle h)
{
    if (!h || !h->comm) return -1;
    return h->p2p_enabled ? 1 : 0;
}

extern "C" int xpcs_comm_nccl_version(void)
{
    NcclApi *api = nccl_api();
    int v = 0;
    if (!api || api->GetVersion(&v) != ncclSuccess) return 0;
    return v;
}
