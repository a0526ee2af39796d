// TEST INFRASTRUCTURE ONLY.  Linked into oracle/_ref/fastgs_standalone_trace (shared cudart): these definitions interpose
// the runtime's cudaMemsetAsync / cudaStreamCreate for the whole executable -- including the calls the UNMODIFIED
// reference objects and the CUB code inlined into them make -- log arguments and results to stderr and forward to the
// real entry points.  Nothing of the reference is changed.
#include <cstdio>
#include <cuda_runtime_api.h>
#include <dlfcn.h>

extern "C" cudaError_t cudaMemsetAsync(void* devPtr, int value, size_t count, cudaStream_t stream) {
    using fn_t = cudaError_t (*)(void*, int, size_t, cudaStream_t);
    static fn_t real = (fn_t)dlsym(RTLD_NEXT, "cudaMemsetAsync");
    cudaError_t r = real(devPtr, value, count, stream);
    cudaPointerAttributes at{};
    cudaError_t ra = cudaPointerGetAttributes(&at, devPtr);
    fprintf(stderr, "[trace] cudaMemsetAsync(ptr=%p, value=%d, count=%zu, stream=%p) -> %s | pointer type=%d (%s)\n", devPtr,
            value, count, (void*)stream, cudaGetErrorName(r), ra == cudaSuccess ? (int)at.type : -1, cudaGetErrorName(ra));
    return r;
}

extern "C" cudaError_t cudaStreamCreate(cudaStream_t* s) {
    using fn_t = cudaError_t (*)(cudaStream_t*);
    static fn_t real = (fn_t)dlsym(RTLD_NEXT, "cudaStreamCreate");
    cudaError_t r = real(s);
    fprintf(stderr, "[trace] cudaStreamCreate -> %s, stream=%p\n", cudaGetErrorName(r), s ? (void*)*s : nullptr);
    return r;
}
