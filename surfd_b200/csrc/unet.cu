// unet.cu -- placeholder until the sampler kernels land (next commit); keeps the C ABI complete.
#include "common.cuh"
using namespace surfd;
struct surfd_unet { int dummy; };
extern "C" size_t surfd_unet_packed_floats(void) { return 0; }
extern "C" int surfd_unet_create(const float*, size_t, int, int, int, surfd_unet**) { return set_error(SURFD_BAD_ARGUMENT, "sampler not built yet", __FILE__, __LINE__); }
extern "C" void surfd_unet_destroy(surfd_unet*) {}
extern "C" int surfd_unet_forward(surfd_unet*, int, const float*, const int64_t*, const float*, const int64_t*, float*, void*) { return set_error(SURFD_BAD_ARGUMENT, "sampler not built yet", __FILE__, __LINE__); }
extern "C" int surfd_sample(surfd_unet*, int, int, const int64_t*, const float*, const float*, const float*, const int64_t*, float, float*, void*) { return set_error(SURFD_BAD_ARGUMENT, "sampler not built yet", __FILE__, __LINE__); }
