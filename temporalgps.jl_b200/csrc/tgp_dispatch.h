// tgp_dispatch.h — declarations of the per-dimension drivers (defined in tgp_drivers.cuh,
// instantiated in tgp_inst.cu once per supported D).
#pragma once
#include "tgp_ctx.cuh"

namespace tgp {

template <int D> int do_filter(tgp_ctx* h, const tgp_lgssm* m, const double* y, double* m_f, int64_t s_m, double* P_f,
                               int64_t s_P, double* lml_out, double* lml_steps);
template <int D> int do_posterior_marginals(tgp_ctx* h, const tgp_lgssm* m, const double* y, const double* R_new,
                                            int64_t sRnew, double* mean_out, double* var_out, double* lml_out);
template <int D> int do_posterior(tgp_ctx* h, const tgp_lgssm* m, const double* y, double* G, double* g, double* Sig,
                                  double* m_T, double* P_T);
template <int D> int do_marginals(tgp_ctx* h, const tgp_lgssm* m, double* mean_out, double* var_out);
template <int D> int do_shard_reduce(tgp_ctx* h, const tgp_lgssm* m, const double* y, double* elem_out);
template <int D> int do_shard_phase1(tgp_ctx* h, const tgp_lgssm* m, const double* y, int rank, int world, double* xchg_out);
template <int D> int do_shard_phase2(tgp_ctx* h, const double* xchg_all, double* lml_partial);
template <int D> int do_shard_logpdf(tgp_ctx* h, const tgp_lgssm* m, const double* y, int rank, int world, bool* handled);
template <int D> int do_shard_prefix(int n, const double* elems, const double* m0, const double* P0, double* m_in,
                                     double* P_in);

// Large-state / vector-observation path (tgp_dense.cu): any D, M; sequential in time.
int dense_filter(tgp_ctx* h, const tgp_lgssm* m, const double* y, double* lml_out, double* lml_steps, double* m_f, int64_t s_m,
                 double* P_f, int64_t s_P);

void dense_release(tgp_ctx* h);   // frees the dense path's cuBLAS handle

// Step-by-step smoother-side path for the shapes the scan kernels do not cover (tgp_seq.cu): vector observations, Reverse models.
int seq_posterior(tgp_ctx* h, const tgp_lgssm* m, const double* y, double* G, double* g, double* Sig, double* m_T, double* P_T, double* lml_out);
int seq_marginals(tgp_ctx* h, const tgp_lgssm* m, double* mean_out, double* cov_out, int diag);
int seq_posterior_marginals(tgp_ctx* h, const tgp_lgssm* m, const double* y, const double* R_new, int64_t sRnew, double* mean_out,
                            double* var_out, double* lml_out);

// Device-side construction of (A[t], Q[t]) of an LTI SDE on an irregular grid (tgp_lti.cu).
int lti_components(tgp_ctx* h, int D, int64_t T, const double* F, const double* F0, const double* P, const double* t, double* A_out,
                   double* Q_out);

// Peer-memory exchange of the time-sharded path (tgp_xchg.cu).
int xchg_create(tgp_ctx* h, int rank, int world, int slot_doubles, void* ipc_handle_out);
int xchg_open(tgp_ctx* h, const void* handles_all);
void xchg_destroy(tgp_ctx* h);
// Device view of an opened exchange (false: none is open) and its per-channel epochs, for the kernels that put / wait themselves.
struct XchgView { char* const* peers; char* self; int slot, world, rank; unsigned long long flag_off; };
bool xchg_view(tgp_ctx* h, XchgView* v);
unsigned long long xchg_next_epoch(tgp_ctx* h, int ch);
unsigned long long xchg_epoch(tgp_ctx* h, int ch);

// Region of the exchange buffer used by the one-launch sharded logpdf (tgp_fir.cuh), identical on every rank:
//   halo[4][3 * 1024] doubles   ring (by call epoch & 3) of the observations that precede this rank's shard, written by rank - 1
//   halo_flag[4], ack           u64: epoch of the halo in each ring slot (written by rank - 1) / of the last halo rank + 1 has consumed
//   lml[4][world] 16-byte words ring of the shards' log-likelihoods: {value, epoch}, one 128-bit store each
struct FirXchgLayout {
    static constexpr int kRing = 4, kHaloDoubles = 3 * 1024;
    static __host__ __device__ constexpr size_t halo_off(unsigned long long epoch) { return (size_t)(epoch & 3ull) * kHaloDoubles * sizeof(double); }
    static __host__ __device__ constexpr size_t halo_flag_off(unsigned long long epoch) { return (size_t)kRing * kHaloDoubles * sizeof(double) + (size_t)(epoch & 3ull) * 8; }
    static __host__ __device__ constexpr size_t ack_off() { return halo_flag_off(0) + 64; }
    static __host__ __device__ constexpr size_t lml_off(int world, unsigned long long epoch, int rank) { return ack_off() + 64 + ((size_t)(epoch & 3ull) * world + rank) * 16; }
    static __host__ __device__ constexpr size_t bytes(int world) { return lml_off(world, 3, world) + 256; }
};
struct XchgFirView { char* const* peers; char* self; char* prev; char* next; size_t fir_off; int world, rank; unsigned long long* epoch; };
bool xchg_fir_view(tgp_ctx* h, XchgFirView* v);
int xchg_fir_total(tgp_ctx* h, unsigned long long epoch, double* dst_dev);

// Test hook for the tcgen05 contraction kernel alone (tgp_dense_tc.cuh).
int tc_gemm_selftest(tgp_ctx* h, int K, int Mx, int N, const float* X, const float* Y, float* C, int symmetric);

#define TGP_DECL_D(Dv)                                                                                               \
    extern template int do_filter<Dv>(tgp_ctx*, const tgp_lgssm*, const double*, double*, int64_t, double*, int64_t, \
                                      double*, double*);                                                             \
    extern template int do_posterior_marginals<Dv>(tgp_ctx*, const tgp_lgssm*, const double*, const double*, int64_t, \
                                                   double*, double*, double*);                                       \
    extern template int do_posterior<Dv>(tgp_ctx*, const tgp_lgssm*, const double*, double*, double*, double*,       \
                                         double*, double*);                                                          \
    extern template int do_marginals<Dv>(tgp_ctx*, const tgp_lgssm*, double*, double*);                              \
    extern template int do_shard_reduce<Dv>(tgp_ctx*, const tgp_lgssm*, const double*, double*);                     \
    extern template int do_shard_phase1<Dv>(tgp_ctx*, const tgp_lgssm*, const double*, int, int, double*);           \
    extern template int do_shard_phase2<Dv>(tgp_ctx*, const double*, double*);                                       \
    extern template int do_shard_logpdf<Dv>(tgp_ctx*, const tgp_lgssm*, const double*, int, int, bool*);             \
    extern template int do_shard_prefix<Dv>(int, const double*, const double*, const double*, double*, double*);

// The set of latent dimensions with kernel instantiations (keep in step with build.py's TGP_DIMS).
#define TGP_FOR_EACH_D(X) X(1) X(2) X(3) X(4) X(5) X(6) X(8) X(10)
TGP_FOR_EACH_D(TGP_DECL_D)

}  // namespace tgp
