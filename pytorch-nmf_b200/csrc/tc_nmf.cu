// placeholder until the tcgen05 path lands
#include "tc_nmf.cuh"
namespace nmfb200 {
struct TcState { int dummy; };
bool tc_shape_supported(int64_t, int64_t, int64_t) { return false; }
int tc_create(TcState**, int, int64_t, int64_t, int64_t, bool) { set_error("tensor-core path not built"); return 1; }
void tc_destroy(TcState*) {}
bool tc_supports_beta(const TcState*, double) { return false; }
bool tc_supports_loss(const TcState*, double) { return false; }
int tc_set_target(TcState*, const float*, int64_t, const float*, cudaStream_t) { return 1; }
void tc_mark_dirty(TcState*, bool, bool) {}
int tc_update_w(TcState*, float*, const float*, double, double, double, double, cudaStream_t) { return 1; }
int tc_update_h(TcState*, const float*, float*, double, double, double, double, cudaStream_t) { return 1; }
int tc_w_partial(TcState*, const float*, const float*, double, float*, cudaStream_t) { return 1; }
int tc_contract_only(TcState*, const float*, const float*, int, double, cudaStream_t) { return 1; }
int tc_loss(TcState*, const float*, const float*, double, double*, cudaStream_t) { return 1; }
}
