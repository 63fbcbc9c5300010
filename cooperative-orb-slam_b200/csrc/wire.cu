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

// ---------------------------------------------------------------------------------------------------------------------
// The message itself.  The reference's lcmKeyFrameInfo (R21/include/lcmKeyFrame/lcmKeyFrameInfo.hpp:24-145) carries, per
// key frame, what the server needs to rebuild a KeyFrame without the image (ORB_SLAM2/Examples/ROS/ORB_SLAM2/src/
// ros_mono.cc:230-544, :2108-2167): N, mvKeys / mvKeysUn (int16-truncated), mvuRight, mvDepth, the descriptors (one float
// per byte), BowVector / FeatureVector, and per feature a map-point flag + world position (lcmKeyFrameMapPoints.hpp:19-27).
// Here a message is one contiguous device buffer for a batch of key frames, compacted to the actual key point counts:
//   int32 header[8]  = {magic, n_frames, total, flags, cap, 0, 0, 0}          flags: 1 u_right/depth, 2 map points, 4 kps_un
//   int32 count[n_frames], int32 offset[n_frames]   (padded to 16 bytes)
//   kps [total] x 28 B (lcm-truncated) | kps_un [total] x 28 B (flag 4) | desc [total] x 32 B
//   u_right [total] f32, depth [total] f32 (flag 1) | map points [total] x {x, y, z, has} 16 B (flag 2)
// Feature / BoW vectors are produced on the receiving side from the descriptors (orbv_transform: the vocabulary is resident
// there), which is what they are a function of.
// ---------------------------------------------------------------------------------------------------------------------
constexpr int kWireMagic = 0x4f52424b;      // "ORBK"

struct WireSections { size_t counts, offsets, kps, kps_un, desc, u_right, depth, mappoints, end; };

__host__ __device__ inline size_t wire_align(size_t v) { return (v + 15) & ~(size_t)15; }
__host__ __device__ inline WireSections wire_sections(int n_frames, size_t total, int flags) {
    WireSections w;
    w.counts = 32; w.offsets = w.counts + (size_t)n_frames * 4;
    w.kps = wire_align(w.offsets + (size_t)n_frames * 4);
    w.kps_un = wire_align(w.kps + total * sizeof(orb_keypoint_t));
    w.desc = wire_align(w.kps_un + ((flags & 4) ? total * sizeof(orb_keypoint_t) : 0));
    w.u_right = wire_align(w.desc + total * 32);
    w.depth = wire_align(w.u_right + ((flags & 1) ? total * 4 : 0));
    w.mappoints = wire_align(w.depth + ((flags & 1) ? total * 4 : 0));
    w.end = wire_align(w.mappoints + ((flags & 2) ? total * 16 : 0));
    return w;
}

__device__ inline orb_keypoint_t wire_quantise(orb_keypoint_t k) {
    k.x = through_int16(k.x); k.y = through_int16(k.y); k.size = through_int16(k.size); k.response = through_int16(k.response);
    k.octave = (int16_t)k.octave; k.class_id = (int16_t)k.class_id;
    return k;
}

// one CTA per key frame; the sections are laid out for the WORST case total = n_frames * cap so that the layout does not
// depend on device-side counts (the consumer reads `total` from the header and only moves / reads the used prefix of a section)
__global__ void __launch_bounds__(256) wire_pack_kernel(const orb_keypoint_t* __restrict__ kps, const orb_keypoint_t* __restrict__ kps_un,
                                                        const uint4* __restrict__ desc, const int* __restrict__ counts, const float* __restrict__ u_right,
                                                        const float* __restrict__ depth, const float4* __restrict__ mappoints, int n_frames, int cap,
                                                        int flags, unsigned char* __restrict__ msg) {
    const int f = blockIdx.x;
    __shared__ int s_off, s_total;
    if (threadIdx.x == 0) {
        int off = 0, tot = 0;
        for (int i = 0; i < n_frames; i++) { const int c = min(max(counts[i], 0), cap); if (i < f) off += c; tot += c; }
        s_off = off; s_total = tot;
    }
    __syncthreads();
    const int off = s_off, n = min(max(counts[f], 0), cap);
    const WireSections w = wire_sections(n_frames, (size_t)n_frames * cap, flags);
    int* hdr = reinterpret_cast<int*>(msg);
    if (threadIdx.x == 0) {
        if (f == 0) { hdr[0] = kWireMagic; hdr[1] = n_frames; hdr[2] = s_total; hdr[3] = flags; hdr[4] = cap; hdr[5] = hdr[6] = hdr[7] = 0; }
        reinterpret_cast<int*>(msg + w.counts)[f] = n;
        reinterpret_cast<int*>(msg + w.offsets)[f] = off;
    }
    orb_keypoint_t* o_k = reinterpret_cast<orb_keypoint_t*>(msg + w.kps) + off;
    orb_keypoint_t* o_ku = reinterpret_cast<orb_keypoint_t*>(msg + w.kps_un) + off;
    uint4* o_d = reinterpret_cast<uint4*>(msg + w.desc) + 2 * (size_t)off;
    float* o_ur = reinterpret_cast<float*>(msg + w.u_right) + off;
    float* o_dp = reinterpret_cast<float*>(msg + w.depth) + off;
    float4* o_mp = reinterpret_cast<float4*>(msg + w.mappoints) + off;
    const size_t base = (size_t)f * cap;
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        o_k[i] = wire_quantise(kps[base + i]);
        if (flags & 4) o_ku[i] = wire_quantise(kps_un[base + i]);
        if (flags & 1) { o_ur[i] = u_right[base + i]; o_dp[i] = depth[base + i]; }
        if (flags & 2) o_mp[i] = mappoints[base + i];
    }
    for (int i = threadIdx.x; i < 2 * n; i += blockDim.x) o_d[i] = desc[2 * base + i];
}

// blockIdx.y = message (one per sending agent, msg_stride bytes apart); the outputs are [n_msgs][n_frames][cap]
__global__ void __launch_bounds__(256) wire_unpack_kernel(const unsigned char* __restrict__ msg, size_t msg_stride, int n_frames, int cap, int flags,
                                                          orb_keypoint_t* __restrict__ kps, orb_keypoint_t* __restrict__ kps_un, uint4* __restrict__ desc,
                                                          int* __restrict__ counts, float* __restrict__ u_right, float* __restrict__ depth,
                                                          float4* __restrict__ mappoints) {
    const int f = blockIdx.x;
    {
        const size_t m = blockIdx.y, fo = m * n_frames * (size_t)cap;
        msg += m * msg_stride; kps += fo; desc += 2 * fo; counts += m * n_frames;
        if (kps_un) kps_un += fo;
        if (u_right) { u_right += fo; depth += fo; }
        if (mappoints) mappoints += fo;
    }
    const int* hdr = reinterpret_cast<const int*>(msg);
    if (hdr[0] != kWireMagic || hdr[1] != n_frames || hdr[4] != cap || hdr[3] != flags) { if (threadIdx.x == 0) counts[f] = -1; return; }
    const WireSections w = wire_sections(n_frames, (size_t)n_frames * cap, flags);
    const int n = reinterpret_cast<const int*>(msg + w.counts)[f], off = reinterpret_cast<const int*>(msg + w.offsets)[f];
    if (threadIdx.x == 0) counts[f] = n;
    const orb_keypoint_t* i_k = reinterpret_cast<const orb_keypoint_t*>(msg + w.kps) + off;
    const orb_keypoint_t* i_ku = reinterpret_cast<const orb_keypoint_t*>(msg + w.kps_un) + off;
    const uint4* i_d = reinterpret_cast<const uint4*>(msg + w.desc) + 2 * (size_t)off;
    const size_t base = (size_t)f * cap;
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        kps[base + i] = i_k[i];
        if ((flags & 4) && kps_un) kps_un[base + i] = i_ku[i];
        if ((flags & 1) && u_right) { u_right[base + i] = reinterpret_cast<const float*>(msg + w.u_right)[off + i]; depth[base + i] = reinterpret_cast<const float*>(msg + w.depth)[off + i]; }
        if ((flags & 2) && mappoints) mappoints[base + i] = reinterpret_cast<const float4*>(msg + w.mappoints)[off + i];
    }
    for (int i = threadIdx.x; i < 2 * n; i += blockDim.x) desc[2 * base + i] = i_d[i];
}

}  // namespace orbcuda

using namespace orbcuda;

extern "C" {

size_t orbw_message_bytes(int n_frames, int cap, int flags) {
    if (n_frames < 0 || cap < 0) return 0;
    return wire_sections(n_frames, (size_t)n_frames * cap, flags).end;
}

int orbw_pack_keyframes_device(const void* d_kps, const void* d_kps_un, const uint8_t* d_desc, const int32_t* d_counts, const float* d_u_right,
                               const float* d_depth, const float* d_mappoints, int n_frames, int cap, void* d_msg, void* stream) {
    if (!d_kps || !d_desc || !d_counts || n_frames < 1 || cap < 1 || !d_msg || (d_u_right && !d_depth) || (reinterpret_cast<uintptr_t>(d_msg) & 15) ||
        (reinterpret_cast<uintptr_t>(d_desc) & 15)) { set_error("orbw_pack_keyframes_device: bad arguments"); return ORB_ERR_ARG; }
    const int flags = (d_u_right ? 1 : 0) | (d_mappoints ? 2 : 0) | (d_kps_un ? 4 : 0);
    wire_pack_kernel<<<n_frames, 256, 0, (cudaStream_t)stream>>>((const orb_keypoint_t*)d_kps, (const orb_keypoint_t*)d_kps_un, (const uint4*)d_desc, d_counts,
                                                                d_u_right, d_depth, (const float4*)d_mappoints, n_frames, cap, flags, (unsigned char*)d_msg);
    ORB_CUDA_TRY(cudaGetLastError());
    return ORB_OK;
}

int orbw_unpack_keyframes_device(const void* d_msg, int n_msgs, size_t msg_stride, int n_frames, int cap, int flags, void* d_kps, void* d_kps_un,
                                 uint8_t* d_desc, int32_t* d_counts, float* d_u_right, float* d_depth, float* d_mappoints, void* stream) {
    if (!d_msg || n_msgs < 1 || (msg_stride & 15) || n_frames < 1 || cap < 1 || !d_kps || !d_desc || !d_counts || (reinterpret_cast<uintptr_t>(d_msg) & 15) ||
        (reinterpret_cast<uintptr_t>(d_desc) & 15) || (d_u_right && !d_depth)) { set_error("orbw_unpack_keyframes_device: bad arguments"); return ORB_ERR_ARG; }
    wire_unpack_kernel<<<dim3(n_frames, n_msgs), 256, 0, (cudaStream_t)stream>>>((const unsigned char*)d_msg, msg_stride, n_frames, cap, flags, (orb_keypoint_t*)d_kps, (orb_keypoint_t*)d_kps_un,
                                                                  (uint4*)d_desc, d_counts, d_u_right, d_depth, (float4*)d_mappoints);
    ORB_CUDA_TRY(cudaGetLastError());
    return ORB_OK;
}

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
