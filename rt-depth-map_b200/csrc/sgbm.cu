// sgbm.cu -- semi-global matching (placeholder until the kernels land).
#include "common.cuh"
namespace rtdm {
size_t sgbm_work_bytes(const SgbmGeom &, size_t *planes, size_t *vol) { if (planes) *planes = 0; if (vol) *vol = 0; return 0; }
int launch_sgbm(const SgbmGeom &, int, PlaneU8, PlaneU8, PlaneS16, SgbmWork, cudaStream_t, int *)
{
    set_error("sgbm: not implemented yet");
    return -RTDM_ENOSYS;
}
}  // namespace rtdm
struct rtdm_sgbm { int dummy; };
extern "C" int rtdm_sgbm_create(rtdm_sgbm **out, const rtdm_params *, int, int, int, int) { if (out) *out = nullptr; rtdm::set_error("sgbm: not implemented yet"); return -RTDM_ENOSYS; }
extern "C" void rtdm_sgbm_destroy(rtdm_sgbm *) {}
extern "C" int rtdm_sgbm_compute(rtdm_sgbm *, const uint8_t *, size_t, const uint8_t *, size_t, int, int, int16_t *, size_t) { return -RTDM_ENOSYS; }
extern "C" int rtdm_sgbm_compute_batch(rtdm_sgbm *, int, const uint8_t *, size_t, size_t, const uint8_t *, size_t, size_t, int, int, int16_t *, size_t, size_t) { return -RTDM_ENOSYS; }
extern "C" int rtdm_sgbm_compute_device(rtdm_sgbm *, int, const uint8_t *, size_t, size_t, const uint8_t *, size_t, size_t, int, int, int16_t *, size_t, size_t, void *) { return -RTDM_ENOSYS; }
extern "C" int rtdm_sgbm_last_launches(const rtdm_sgbm *) { return 0; }
