// tgp_xchg.cu — peer-memory exchange buffers of the time-sharded path: the halo of a shard and the partial
// log-likelihoods travel by direct NVLink / NVSwitch stores from inside k_fir_logpdf into the peers' buffers (FirXchgLayout), instead
// of an NCCL collective. One process per GPU: the buffers are shared through CUDA IPC handles that the host exchanges once (any
// transport; sharded.py uses torch.distributed). This file owns the buffers and the on-demand sum of the partial results.
#include "tgp_ctx.cuh"
#include "tgp_dispatch.h"

namespace tgp {

constexpr int kXchgChannels = 2;

struct XchgState {
    int rank = 0, world = 0, slot = 0;          // slot: doubles per (channel, parity, rank)
    char* self = nullptr;                       // this rank's buffer (cudaMalloc)
    std::vector<char*> peers;                   // mapped buffers of all ranks (peers[rank] = self)
    char** d_peers = nullptr;                   // the same on the device
    unsigned long long epoch[kXchgChannels] = {0, 0};
    size_t flag_off = 0, bytes = 0;
    size_t fir_off = 0;                         // region of the one-launch sharded logpdf (tgp_fir.cuh), see FirXchgLayout
    unsigned long long fir_epoch = 0;
};

__device__ __forceinline__ size_t xchg_data_off(int ch, int parity, int world, int slot, int r) {
    return ((size_t)((ch * 2 + parity) * world + r) * slot) * sizeof(double);
}


int xchg_create(tgp_ctx* h, int rank, int world, int slot_doubles, void* ipc_handle_out) {
    if (h->xchg) return fail(h, TGP_EINVAL, "tgp_xchg_create called twice on this handle");
    if (world < 1 || world > 128 || rank < 0 || rank >= world || slot_doubles < 1) return fail(h, TGP_EINVAL, "bad rank / world / slot size");
    XchgState* x = new XchgState();
    x->rank = rank; x->world = world; x->slot = slot_doubles;
    x->flag_off = (size_t)kXchgChannels * 2 * world * slot_doubles * sizeof(double);
    x->fir_off = (x->flag_off + (size_t)kXchgChannels * world * sizeof(unsigned long long) + 255) & ~size_t(255);
    x->bytes = x->fir_off + FirXchgLayout::bytes(world);
    if (cudaMalloc((void**)&x->self, x->bytes) != cudaSuccess || cudaMemset(x->self, 0, x->bytes) != cudaSuccess) {
        delete x;
        return fail(h, TGP_ENOMEM, "exchange buffer allocation failed");
    }
    cudaIpcMemHandle_t hd;
    cudaError_t e = cudaIpcGetMemHandle(&hd, x->self);
    if (e != cudaSuccess) {
        cudaFree(x->self);
        delete x;
        return fail(h, TGP_ECUDA, "cudaIpcGetMemHandle failed: %s", cudaGetErrorString(e));
    }
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
    memcpy(ipc_handle_out, &hd, 64);
    h->xchg = x;
    return TGP_OK;
}

int xchg_open(tgp_ctx* h, const void* handles_all) {
    XchgState* x = (XchgState*)h->xchg;
    if (!x) return fail(h, TGP_EINVAL, "tgp_xchg_open before tgp_xchg_create");
    x->peers.assign(x->world, nullptr);
    for (int p = 0; p < x->world; ++p) {
        if (p == x->rank) { x->peers[p] = x->self; continue; }
        cudaIpcMemHandle_t hd;
        memcpy(&hd, (const char*)handles_all + (size_t)p * 64, 64);
        void* ptr = nullptr;
        cudaError_t e = cudaIpcOpenMemHandle(&ptr, hd, cudaIpcMemLazyEnablePeerAccess);
        if (e != cudaSuccess) return fail(h, TGP_ECUDA, "cudaIpcOpenMemHandle(rank %d) failed: %s", p, cudaGetErrorString(e));
        x->peers[p] = (char*)ptr;
    }
    TGP_CUDA(h, cudaMalloc((void**)&x->d_peers, sizeof(char*) * x->world));
    TGP_CUDA(h, cudaMemcpy(x->d_peers, x->peers.data(), sizeof(char*) * x->world, cudaMemcpyHostToDevice));
    return TGP_OK;
}


// Sum of the shards' log-likelihoods of the sharded call with epoch `epoch` (ring slot epoch & 3): waits for every rank's flag.
__global__ void __launch_bounds__(128) k_fir_lml_total(const char* __restrict__ self, size_t fir_off, int world, unsigned long long epoch,
                                                       double* __restrict__ dst) {
    __shared__ double part[128];
    const char* base = self + fir_off;
    if (threadIdx.x < world) {      // {value, epoch}: written with one 128-bit store by rank threadIdx.x's kernel
        const double2* w = reinterpret_cast<const double2*>(base + FirXchgLayout::lml_off(world, epoch, threadIdx.x));
        double v, t;
        unsigned long long spins = 0;
        for (;;) {
            asm volatile("ld.volatile.global.v2.f64 {%0, %1}, [%2];" : "=d"(v), "=d"(t) : "l"(w) : "memory");
            if ((unsigned long long)__double_as_longlong(t) == epoch) break;
            if (++spins > (1ull << 31)) __trap();
            __nanosleep(50);
        }
        part[threadIdx.x] = v;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        double s = 0.0;
        for (int p = 0; p < world; ++p) s += part[p];     // fixed order: the same bits on every rank
        *dst = s;
    }
}

int xchg_fir_total(tgp_ctx* h, unsigned long long epoch, double* dst_dev) {
    XchgState* x = (XchgState*)h->xchg;
    if (!x || !x->d_peers) return fail(h, TGP_EINVAL, "exchange not opened");
    TGP_K(h, "k_fir_lml_total");
    k_fir_lml_total<<<1, 128, 0, h->stream>>>(x->self, x->fir_off, x->world, epoch, dst_dev);
    TGP_LAUNCH_CHECK(h);
    return TGP_OK;
}

bool xchg_fir_view(tgp_ctx* h, XchgFirView* v) {
    XchgState* x = (XchgState*)h->xchg;
    if (!x || !x->d_peers) return false;
    v->peers = x->d_peers;
    v->self = x->self;
    v->prev = x->rank > 0 ? x->peers[x->rank - 1] : nullptr;
    v->next = x->rank + 1 < x->world ? x->peers[x->rank + 1] : nullptr;
    v->fir_off = x->fir_off;
    v->world = x->world;
    v->rank = x->rank;
    v->epoch = &x->fir_epoch;
    return true;
}

bool xchg_view(tgp_ctx* h, XchgView* v) {
    XchgState* x = (XchgState*)h->xchg;
    if (!x || !x->d_peers) return false;
    *v = XchgView{x->d_peers, x->self, x->slot, x->world, x->rank, (unsigned long long)x->flag_off};
    return true;
}
unsigned long long xchg_next_epoch(tgp_ctx* h, int ch) { return ++((XchgState*)h->xchg)->epoch[ch]; }
unsigned long long xchg_epoch(tgp_ctx* h, int ch) { return ((XchgState*)h->xchg)->epoch[ch]; }

void xchg_destroy(tgp_ctx* h) {
    XchgState* x = (XchgState*)h->xchg;
    if (!x) return;
    for (int p = 0; p < (int)x->peers.size(); ++p)
        if (p != x->rank && x->peers[p]) cudaIpcCloseMemHandle(x->peers[p]);
    if (x->d_peers) cudaFree(x->d_peers);
    if (x->self) cudaFree(x->self);
    delete x;
    h->xchg = nullptr;
}

}  // namespace tgp
