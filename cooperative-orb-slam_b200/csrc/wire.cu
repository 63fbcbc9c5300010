// wire.cu -- the agent -> server key-frame message.
// The reference serialises a key frame into an LCM message (R21/Examples/ROS/ORB_SLAM2/src/ros_mono.cc:1929-2399,
// decoded in ORB_SLAM2/Examples/ROS/ORB_SLAM2/src/ros_mono.cc:230-544): key points as lcmKeyPoint
// (include/lcmKeyFrame/lcmKeyPoint.hpp:19-31: int16 x, y, size, response, octave, class_id; float angle), i.e. the float
// members are TRUNCATED to 16-bit integers by the assignment at ros_mono.cc:2071-2077, and the descriptors as one
// float per byte (lcmKeyFrameInfo.hpp:107, filled :2152-2169, decoded back to uchar :402-407), which is lossless.
// On one 8xB200 box the message is the extractor's device-resident output itself (orbx_extract_batch_device: 28-byte
// key points + 32-byte descriptors + counts), exchanged with ncclAllGather; the only semantic step of the wire is the
// int16 truncation, reproduced here so that the receiving side sees exactly what the reference's server sees.
#include <cuda_runtime.h>

#include <cstdint>

#include "internal.h"
#include "orbcuda.h"

namespace orbcuda {

__host__ __device__ inline float through_int16(float v) {
    // float -> int16_t as the C++ conversion at ros_mono.cc:2071-2077 does it (truncation toward zero; values outside
    // int16 are undefined behaviour in the reference -- x86 cvttss2si + 16-bit store wraps; coordinates never get there)
    return (float)(int16_t)(int)v;
}

__global__ void wire_quantize_kernel(orb_keypoint_t* __restrict__ kps, const int* __restrict__ counts, int cap) {
    const int frame = blockIdx.y;
    const int n = counts ? min(counts[frame], cap) : cap;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        orb_keypoint_t k = kps[(size_t)frame * cap + i];
        k.x = through_int16(k.x); k.y = through_int16(k.y); k.size = through_int16(k.size); k.response = through_int16(k.response);
        k.octave = (int16_t)k.octave; k.class_id = (int16_t)k.class_id;
        kps[(size_t)frame * cap + i] = k;
    }
}

}  // namespace orbcuda

using namespace orbcuda;

extern "C" {

int orbw_quantize_lcm_host(orb_keypoint_t* kps, int n) {
    if (n < 0 || (n && !kps)) { set_error("orbw_quantize_lcm_host: bad arguments"); return ORB_ERR_ARG; }
    for (int i = 0; i < n; i++) {
        kps[i].x = through_int16(kps[i].x); kps[i].y = through_int16(kps[i].y);
        kps[i].size = through_int16(kps[i].size); kps[i].response = through_int16(kps[i].response);
        kps[i].octave = (int16_t)kps[i].octave; kps[i].class_id = (int16_t)kps[i].class_id;
    }
    return ORB_OK;
}

int orbw_quantize_lcm_device(void* d_kps, const int32_t* d_counts, int n_frames, int cap, void* stream) {
    if (!d_kps || n_frames < 0 || cap < 0) { set_error("orbw_quantize_lcm_device: bad arguments"); return ORB_ERR_ARG; }
    if (n_frames == 0 || cap == 0) return ORB_OK;
    wire_quantize_kernel<<<dim3((cap + 255) / 256, n_frames), 256, 0, (cudaStream_t)stream>>>((orb_keypoint_t*)d_kps, d_counts, cap);
    ORB_CUDA_TRY(cudaGetLastError());
    return ORB_OK;
}

}  // extern "C"
