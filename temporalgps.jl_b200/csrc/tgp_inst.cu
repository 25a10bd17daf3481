// tgp_inst.cu — explicit instantiation of the per-dimension drivers for ONE latent dimension
// (compile with -DTGP_D=<D>; build.py compiles one object per supported D in parallel).
#ifndef TGP_D
#error "compile with -DTGP_D=<latent dimension>"
#endif
#include "tgp_drivers.cuh"

namespace tgp {
template int do_filter<TGP_D>(tgp_ctx*, const tgp_lgssm*, const double*, double*, int64_t, double*, int64_t, double*, double*);
template int do_posterior_marginals<TGP_D>(tgp_ctx*, const tgp_lgssm*, const double*, const double*, int64_t, double*, double*,
                                           double*);
template int do_posterior<TGP_D>(tgp_ctx*, const tgp_lgssm*, const double*, double*, double*, double*, double*, double*);
template int do_marginals<TGP_D>(tgp_ctx*, const tgp_lgssm*, double*, double*);
template int do_shard_reduce<TGP_D>(tgp_ctx*, const tgp_lgssm*, const double*, double*);
template int do_shard_phase1<TGP_D>(tgp_ctx*, const tgp_lgssm*, const double*, int, int, double*);
template int do_shard_phase2<TGP_D>(tgp_ctx*, const double*, double*);
template int do_shard_logpdf<TGP_D>(tgp_ctx*, const tgp_lgssm*, const double*, int, int, bool*);
template int do_shard_prefix<TGP_D>(int, const double*, const double*, const double*, double*, double*);
}  // namespace tgp
