// placeholder - replaced by the spectral solver
#include "common.cuh"
using namespace tq;
extern "C" int tq_solver_workspace(int64_t n, size_t* bytes) { *bytes = 256; return TQ_OK; }
extern "C" int tq_spectral_solve(const double*, int64_t, int64_t, double, int, double*, double*, int64_t*, double*, int64_t*, void*, size_t, void*) { set_error("not implemented"); return TQ_ERR_UNSUPPORTED; }
extern "C" int tq_eigh(const double*, int64_t, int64_t, double*, double*, int64_t, void*, size_t, void*) { set_error("not implemented"); return TQ_ERR_UNSUPPORTED; }
extern "C" int tq_rank_select(const double*, int64_t, double, int, double*, int64_t*, void*, size_t, void*) { set_error("not implemented"); return TQ_ERR_UNSUPPORTED; }
extern "C" int tq_qrcp(const double*, int64_t, int64_t, int64_t, double*, int64_t, int64_t*, void*, size_t, void*) { set_error("not implemented"); return TQ_ERR_UNSUPPORTED; }
extern "C" int tq_qr_r(const double*, int64_t, int64_t, int64_t, double*, int64_t, void*, size_t, void*) { set_error("not implemented"); return TQ_ERR_UNSUPPORTED; }
