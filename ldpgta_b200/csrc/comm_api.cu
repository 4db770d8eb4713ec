// comm_api.cu -- the one cross-GPU exchange of the path: the sum of the ranks' chromosome partial sums
// (mean_and_std_of_mean_of_rnd_var needs sums over ALL windows of a chromosome, ANEUPLOIDY_TEST.py:65-76), done by ONE
// kernel over NVLink / NVSwitch peer memory instead of a library collective.  The payload is a few tens of kilobytes, so
// the exchange is pure latency, and it is a PUSH: every rank stores its partials into its slot of every peer's buffer
// (posted remote stores), raises a flag in every peer's memory, waits for the peers' flags in its own memory (local
// polling) and then adds the world's slots from its own HBM in rank order -- every rank obtains bit-identical sums, and
// no load crosses NVLink (the first version pulled the peers' partials with peer loads: one more round trip per pass).
//
// One process per GPU: buffers are shared through CUDA IPC handles, exchanged by the caller (torch.distributed
// all_gather_object in ldpgta_b200/sharding.py, any byte transport will do).  Contexts of one process (tests) use the
// pointers directly with peer access enabled.
#include <unistd.h>
#include <algorithm>

#include "common.cuh"

namespace ldpgta {

struct CommHandle {                      // what ldpgta_comm_create hands out (fits LDPGTA_COMM_HANDLE_BYTES)
    cudaIpcMemHandle_t ipc;              // 64 bytes
    int64_t pid;
    uint64_t ptr;                        // the allocation's address in the owning process
    int32_t device, rank;
    int64_t max_doubles;
};
static_assert(sizeof(CommHandle) <= LDPGTA_COMM_HANDLE_BYTES, "handle must fit the ABI's byte array");

struct PeerComm {
    int rank = 0, world = 1;
    int64_t max_doubles = 0;
    unsigned char *mine = nullptr;       // own allocation: [2][world][max_doubles] doubles, then [2][MAXPEERS] arrival epochs
    unsigned char *peer[LDPGTA_MAX_PEERS] = {};
    bool opened[LDPGTA_MAX_PEERS] = {};
    bool connected = false;
    uint64_t epoch = 0;
    unsigned int *h_status = nullptr;    // page-locked, device-mapped status word (written by the kernel on a time-out)
    unsigned int *d_status = nullptr;    // its device address
};

struct CommView {
    double *data[LDPGTA_MAX_PEERS];              // peers' [2][world][max_doubles]: slot [parity][r] is written by rank r
    unsigned long long *arrive[LDPGTA_MAX_PEERS];   // peers' [2][MAXPEERS]
    unsigned int *status;                        // own: 1 = a peer did not arrive in time
    int rank, world;
    int64_t max_doubles;
};

static size_t comm_data_bytes(int world, int64_t max_doubles) { return (size_t)2 * world * max_doubles * 8; }
static size_t comm_bytes(int world, int64_t max_doubles) { return comm_data_bytes(world, max_doubles) + 2 * LDPGTA_MAX_PEERS * 8 + 64; }

// partials (in place) = sum over ranks, rank order.  One CTA.
static __global__ void __launch_bounds__(1024) k_allreduce_peer(const CommView v, double *__restrict__ partials, int64_t n, unsigned long long epoch)
{
    const int par = (int)(epoch & 1ull);
    const int64_t slot = ((int64_t)par * v.world + v.rank) * v.max_doubles;      // my slot, the same offset in every buffer
    for (int64_t i = threadIdx.x; i < n; i += blockDim.x) {
        const double x = partials[i];
#pragma unroll
        for (int r = 0; r < LDPGTA_MAX_PEERS; ++r)
            if (r < v.world) v.data[r][slot + i] = x;
    }
    __syncthreads();                                   // the block's stores are ordered before the fence of the flag writers
    if ((int)threadIdx.x < v.world) {
        // raise my flag in peer t's memory, then wait for peer t's flag in mine
        const int t = threadIdx.x;
        __threadfence_system();
        volatile unsigned long long *theirs = v.arrive[t] + par * LDPGTA_MAX_PEERS + v.rank;
        *theirs = epoch;
        volatile unsigned long long *mine = v.arrive[v.rank] + par * LDPGTA_MAX_PEERS + t;
        const long long t0 = clock64();
        while (*mine < epoch) {
            if (clock64() - t0 > 4000000000ll) {                      // ~2 s: a peer is gone; report instead of hanging the GPU
                *v.status = 1u;                                       // page-locked host memory, read by the next call
                break;
            }
        }
        __threadfence_system();
    }
    __syncthreads();
    const double *mine = v.data[v.rank] + (int64_t)par * v.world * v.max_doubles;
    for (int64_t i = threadIdx.x; i < n; i += blockDim.x) {
        double s = 0.0;
        for (int r = 0; r < v.world; ++r) s += __ldcv(mine + (int64_t)r * v.max_doubles + i);     // written by the peers: not through L1
        partials[i] = s;
    }
}

static void comm_close(ldpgta_ctx *c)
{
    PeerComm *pc = (PeerComm *)c->comm;
    if (!pc) return;
    if (c->stream) cudaStreamSynchronize(c->stream);
    for (int r = 0; r < pc->world; ++r)
        if (pc->opened[r] && pc->peer[r]) cudaIpcCloseMemHandle(pc->peer[r]);
    if (pc->mine) cudaFree(pc->mine);
    if (pc->h_status) cudaFreeHost(pc->h_status);
    delete pc;
    c->comm = nullptr;
}

void free_comm(ldpgta_ctx *c) { comm_close(c); }

}  // namespace ldpgta

using namespace ldpgta;

extern "C" {

int ldpgta_comm_create(ldpgta_ctx *c, int rank, int world, int64_t max_doubles, unsigned char *handle_out)
{
    if (!c || !handle_out) return fail(LDPGTA_E_INVALID, "null argument");
    if (world < 1 || world > LDPGTA_MAX_PEERS || rank < 0 || rank >= world)
        return fail(LDPGTA_E_INVALID, "rank %d of %d: at most %d peers are supported", rank, world, LDPGTA_MAX_PEERS);
    if (max_doubles < 1) return fail(LDPGTA_E_INVALID, "max_doubles must be positive");
    LDPGTA_CUDA(cudaSetDevice(c->device));
    comm_close(c);
    PeerComm *pc = new PeerComm();
    c->comm = pc;
    pc->rank = rank; pc->world = world; pc->max_doubles = max_doubles;
    LDPGTA_CUDA(cudaMalloc((void **)&pc->mine, comm_bytes(world, max_doubles)));
    LDPGTA_CUDA(cudaMemset(pc->mine, 0, comm_bytes(world, max_doubles)));
    LDPGTA_CUDA(cudaHostAlloc((void **)&pc->h_status, 16, cudaHostAllocMapped));
    *pc->h_status = 0;
    LDPGTA_CUDA(cudaHostGetDevicePointer((void **)&pc->d_status, pc->h_status, 0));
    LDPGTA_CUDA(cudaDeviceSynchronize());
    CommHandle h;
    memset(&h, 0, sizeof h);
    LDPGTA_CUDA(cudaIpcGetMemHandle(&h.ipc, pc->mine));
    h.pid = (int64_t)getpid(); h.ptr = (uint64_t)(uintptr_t)pc->mine; h.device = c->device; h.rank = rank; h.max_doubles = max_doubles;
    memset(handle_out, 0, LDPGTA_COMM_HANDLE_BYTES);
    memcpy(handle_out, &h, sizeof h);
    return 0;
}

int ldpgta_comm_connect(ldpgta_ctx *c, const unsigned char *handles)
{
    if (!c || !handles) return fail(LDPGTA_E_INVALID, "null argument");
    PeerComm *pc = (PeerComm *)c->comm;
    if (!pc) return fail(LDPGTA_E_STATE, "call ldpgta_comm_create first");
    LDPGTA_CUDA(cudaSetDevice(c->device));
    for (int r = 0; r < pc->world; ++r) {
        CommHandle h;
        memcpy(&h, handles + (size_t)r * LDPGTA_COMM_HANDLE_BYTES, sizeof h);
        if (h.rank != r || h.max_doubles != pc->max_doubles)
            return fail(LDPGTA_E_INVALID, "handle %d belongs to rank %d with %lld doubles (expected %lld)", r, h.rank, (long long)h.max_doubles,
                        (long long)pc->max_doubles);
        if (r == pc->rank) { pc->peer[r] = pc->mine; continue; }
        if (h.pid == (int64_t)getpid()) {
            // another context of this process: the pointer is valid here, peer access makes it reachable
            if (h.device != c->device) {
                int can = 0;
                LDPGTA_CUDA(cudaDeviceCanAccessPeer(&can, c->device, h.device));
                if (!can) return fail(LDPGTA_E_UNSUPPORTED, "device %d cannot access device %d's memory", c->device, h.device);
                const cudaError_t e = cudaDeviceEnablePeerAccess(h.device, 0);
                if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) LDPGTA_CUDA(e);
                cudaGetLastError();
            }
            pc->peer[r] = (unsigned char *)(uintptr_t)h.ptr;
        } else {
            void *p = nullptr;
            LDPGTA_CUDA(cudaIpcOpenMemHandle(&p, h.ipc, cudaIpcMemLazyEnablePeerAccess));
            pc->peer[r] = (unsigned char *)p;
            pc->opened[r] = true;
        }
    }
    pc->connected = true;
    return 0;
}

int ldpgta_allreduce_chrom_dev(ldpgta_ctx *c, double *partials_dev, int64_t n_doubles)
{
    if (!c || !partials_dev || n_doubles < 0) return fail(LDPGTA_E_INVALID, "bad arguments");
    PeerComm *pc = (PeerComm *)c->comm;
    if (!pc || !pc->connected) return fail(LDPGTA_E_STATE, "no peer communicator: ldpgta_comm_create + ldpgta_comm_connect first");
    if (n_doubles > pc->max_doubles) return fail(LDPGTA_E_INVALID, "%lld doubles exceed the communicator's %lld", (long long)n_doubles, (long long)pc->max_doubles);
    if (*pc->h_status) return fail(LDPGTA_E_CUDA, "an earlier all-reduce timed out waiting for a peer");
    LDPGTA_CUDA(cudaSetDevice(c->device));
    if (pc->world == 1 || n_doubles == 0) return 0;
    CommView v;
    memset(&v, 0, sizeof v);
    for (int r = 0; r < pc->world; ++r) {
        v.data[r] = (double *)pc->peer[r];
        v.arrive[r] = (unsigned long long *)(pc->peer[r] + comm_data_bytes(pc->world, pc->max_doubles));
    }
    v.status = pc->d_status;
    v.rank = pc->rank; v.world = pc->world; v.max_doubles = pc->max_doubles;
    pc->epoch++;
    const int threads = (int)std::min<int64_t>(1024, std::max<int64_t>(64, (n_doubles + 31) / 32 * 32));
    k_allreduce_peer<<<1, threads, 0, c->stream>>>(v, partials_dev, n_doubles, (unsigned long long)pc->epoch);
    c->launches++;
    LDPGTA_CUDA(cudaGetLastError());
    return 0;
}

int ldpgta_comm_destroy(ldpgta_ctx *c)
{
    if (!c) return 0;
    cudaSetDevice(c->device);
    comm_close(c);
    return 0;
}

}  // extern "C"
