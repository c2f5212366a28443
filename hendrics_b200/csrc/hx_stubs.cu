// TEMPORARY: entry points not implemented yet (replaced by hx_qffa.cu / hx_zn.cu).
#include "hx_common.cuh"
extern "C" {
int hx_qffa_search(const double *, int64_t, const double *, const double *, const double *, int64_t, int, int,
                   const double *, const double *, int64_t, int, double *, uint32_t *, int) {
    hx::set_error("hx_qffa_search: not implemented");
    return HX_ERR_INTERNAL;
}
int hx_qffa_search_dev(const double *, int64_t, double, double, const double *, const double *, const double *,
                       int64_t, int, int, const double *, const double *, int64_t, int, double *, uint32_t *, void *) {
    hx::set_error("hx_qffa_search_dev: not implemented");
    return HX_ERR_INTERNAL;
}
int hx_zn_unbinned(const double *, int64_t, const double *, int64_t, const double *, int64_t, int, double *, int) {
    hx::set_error("hx_zn_unbinned: not implemented");
    return HX_ERR_INTERNAL;
}
int hx_zn_unbinned_dev(const double *, int64_t, const double *, int64_t, const double *, int64_t, int, double *,
                       void *) {
    hx::set_error("hx_zn_unbinned_dev: not implemented");
    return HX_ERR_INTERNAL;
}
}
