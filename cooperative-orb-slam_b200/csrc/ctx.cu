// ctx.cu -- per-thread matcher workspace (see MatchCtx in internal.h).
#include "internal.h"

#include <cstring>

namespace orbcuda {

static inline size_t up256(size_t v) { return (v + 255) & ~(size_t)255; }

MatchCtx& match_ctx() {
    static thread_local MatchCtx ctx;
    return ctx;
}

MatchCtx::~MatchCtx() {
    // best effort: the CUDA runtime may already be shutting down when a thread exits
    if (device >= 0 && cudaSetDevice(device) == cudaSuccess) {
        if (dbase) cudaFree(dbase);
        if (hbase) cudaFreeHost(hbase);
        if (stream) cudaStreamDestroy(stream);
    }
    cudaGetLastError();
}

bool MatchCtx::begin(int dev, size_t dev_bytes, size_t host_bytes) {
    if (cudaSetDevice(dev) != cudaSuccess) {
        cudaGetLastError();
        set_error("no usable CUDA device %d (this library has no CPU fallback)", dev);
        return false;
    }
    if (dev != device) {
        if (device >= 0) {   // the thread moved to another device: drop the old workspace
            cudaSetDevice(device);
            if (dbase) cudaFree(dbase);
            if (hbase) cudaFreeHost(hbase);
            if (stream) cudaStreamDestroy(stream);
            cudaSetDevice(dev);
        }
        dbase = nullptr; hbase = nullptr; stream = nullptr; dcap = hcap = 0;
        device = dev;
        if (!cuda_ok(cudaStreamCreateWithFlags(&stream, cudaStreamNonBlocking), "cudaStreamCreate")) { device = -1; return false; }
    }
    dev_bytes = up256(dev_bytes) + 4096; host_bytes = up256(host_bytes) + 4096;
    if (dev_bytes > dcap) {
        if (!cuda_ok(cudaStreamSynchronize(stream), "cudaStreamSynchronize")) return false;
        if (dbase) cudaFree(dbase);
        dbase = nullptr; dcap = 0;
        const size_t want = dev_bytes + dev_bytes / 2;
        if (!cuda_ok(cudaMalloc((void**)&dbase, want), "cudaMalloc")) return false;
        dcap = want;
    }
    if (host_bytes > hcap) {
        if (!cuda_ok(cudaStreamSynchronize(stream), "cudaStreamSynchronize")) return false;
        if (hbase) cudaFreeHost(hbase);
        hbase = nullptr; hcap = 0;
        const size_t want = host_bytes + host_bytes / 2;
        if (!cuda_ok(cudaHostAlloc((void**)&hbase, want, cudaHostAllocDefault), "cudaHostAlloc")) return false;
        hcap = want;
    }
    doff = hoff = 0; npend = 0;
    up_bytes = dl_bytes = 0; copy_failed = false;
    return true;
}

cudaStream_t MatchCtx::s() {
    if (up_bytes) {
        if (!cuda_ok(cudaMemcpyAsync(up_d, up_h, up_bytes, cudaMemcpyHostToDevice, stream), "cudaMemcpyAsync")) copy_failed = true;
        up_bytes = 0;
    }
    return stream;
}

void* MatchCtx::dalloc(size_t bytes) {
    bytes = up256(bytes ? bytes : 1);
    if (doff + bytes > dcap) { set_error("matcher workspace overflow (device)"); return nullptr; }
    void* p = dbase + doff;
    doff += bytes;
    return p;
}

void* MatchCtx::upload(const void* src, size_t bytes) {
    void* d = dalloc(bytes);
    if (!d) return nullptr;
    if (bytes == 0) return d;
    const size_t hb = up256(bytes);
    if (hoff + hb > hcap) { set_error("matcher workspace overflow (host)"); return nullptr; }
    void* hstage = hbase + hoff;
    hoff += hb;
    memcpy(hstage, src, bytes);
    if (up_bytes && (char*)d == up_d + up256(up_bytes) && (char*)hstage == up_h + up256(up_bytes)) {
        up_bytes = up256(up_bytes) + bytes;      // the padding in between is arena space of nobody
    } else {
        s();
        up_d = (char*)d; up_h = (char*)hstage; up_bytes = bytes;
    }
    return copy_failed ? nullptr : d;
}

bool MatchCtx::download(void* dst, const void* dsrc, size_t bytes) {
    if (bytes == 0) return true;
    const size_t hb = up256(bytes);
    if (hoff + hb > hcap || npend >= 8) { set_error("matcher workspace overflow (download)"); return false; }
    char* hstage = hbase + hoff;
    hoff += hb;
    if (dl_bytes && (const char*)dsrc == dl_d + up256(dl_bytes) && hstage == dl_h + up256(dl_bytes)) {
        dl_bytes = up256(dl_bytes) + bytes;
    } else {
        if (dl_bytes && !cuda_ok(cudaMemcpyAsync(dl_h, dl_d, dl_bytes, cudaMemcpyDeviceToHost, s()), "cudaMemcpyAsync")) return false;
        dl_d = (const char*)dsrc; dl_h = hstage; dl_bytes = bytes;
    }
    pend[npend++] = Pending{dst, hstage, bytes};
    return true;
}

bool MatchCtx::finish() {
    cudaStream_t st = s();
    if (dl_bytes && !cuda_ok(cudaMemcpyAsync(dl_h, dl_d, dl_bytes, cudaMemcpyDeviceToHost, st), "cudaMemcpyAsync")) copy_failed = true;
    dl_bytes = 0;
    if (!cuda_ok(cudaStreamSynchronize(stream), "matcher kernel") || copy_failed) return false;
    for (int i = 0; i < npend; i++) memcpy(pend[i].dst, pend[i].staged, pend[i].bytes);
    npend = 0;
    return true;
}

}  // namespace orbcuda
