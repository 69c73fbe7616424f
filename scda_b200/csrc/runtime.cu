// Small runtime queries for the host side.
#include "common.cuh"

// Identity of the stream capture `stream` currently belongs to (0: not capturing).  The host side keeps lazily
// derived tensors (concatenated / padded weight copies) that several streams share; a consumer may wait on the
// builder's event only inside the SAME capture (or outside any capture) — waiting on an event of another
// capture invalidates the one in progress.
SCDA_API unsigned long long scda_stream_capture_id(cudaStream_t stream)
{
    cudaStreamCaptureStatus status = cudaStreamCaptureStatusNone;
    unsigned long long id = 0;
    if (cudaStreamGetCaptureInfo(stream, &status, &id) != cudaSuccess) {
        (void)cudaGetLastError();
        return 0;
    }
    return status == cudaStreamCaptureStatusActive ? id : 0ull;
}

// A device-side clock reading (%globaltimer, ns) written to buf[slot] in stream order: phase boundaries of a
// replayed iteration graph measured from inside the graph, without a profiler attached.
namespace {
__global__ void timestamp_kernel(unsigned long long *buf, int slot)
{
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    buf[slot] = t;
}
}  // namespace

SCDA_API int scda_timestamp(unsigned long long *buf, int slot, cudaStream_t stream)
{
    if (!buf || slot < 0) return 0;
    timestamp_kernel<<<1, 1, 0, stream>>>(buf, slot);
    return scda_launch_status();
}
