// tgp_steady.cuh — steady-state fast path for time-invariant models (placeholder: not yet enabled).
#pragma once
#include "tgp_ctx.cuh"
#include "tgp_scan_small.cuh"

namespace tgp {

template <int D>
int filter_steady(tgp_ctx* h, const tgp_lgssm& d, const double* dy, FilterReq& rq, bool* handled) {
    (void)h; (void)d; (void)dy; (void)rq;
    *handled = false;
    return TGP_OK;
}

}  // namespace tgp
