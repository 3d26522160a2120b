// clhandler.cu -- CLHandler.h over the CUDA runtime, plus the error convention.
//
// The reference's runtime layer (src/CLHandler.c) wraps platform / device /
// context / program / queue / kernel / buffer / enqueue of OpenCL and aborts on
// any error (src/error.c:147-154).  The same eight steps, CUDA-shaped.
#include <cuda_runtime.h>

#include <cstdio>
#include <cstdlib>
#include <cstring>

#include "CLHandler.h"
#include "clpt_device.cuh"

// clstate.cu
void clpt_state_launch_frame(int width, int height);
cudaStream_t clpt_state_stream();
int clpt_state_device();

struct CLPlatform_ {
    int driver_version, runtime_version, device_count;
};

namespace {
CLPlatform_ g_platform;
bool verbose() {
    const char *v = getenv("CLPT_VERBOSE");
    return v && *v && strcmp(v, "0") != 0;
}
} // namespace

extern "C" {

const char *err_string(int error) { return cudaGetErrorName((cudaError_t)error); }

void handle_err(int err, const char *file, int line) {
    if (err == (int)cudaSuccess) return;
    fprintf(stderr, "%s:%d: CUDA Error: %s (%s)\n", file, line, err_string(err),
            cudaGetErrorString((cudaError_t)err));
    exit(EXIT_FAILURE);
}

CLPlatform CLGetPlatform(void) {
    HANDLE_ERR(cudaDriverGetVersion(&g_platform.driver_version));
    HANDLE_ERR(cudaRuntimeGetVersion(&g_platform.runtime_version));
    HANDLE_ERR(cudaGetDeviceCount(&g_platform.device_count));
    if (g_platform.device_count < 1) {
        fprintf(stderr, "%s:%d: CUDA Error: no CUDA device (this library has no CPU path)\n", __FILE__, __LINE__);
        exit(EXIT_FAILURE);
    }
    if (verbose()) {
        printf("Platform: CUDA driver %d runtime %d, %d device(s)\n", g_platform.driver_version,
               g_platform.runtime_version, g_platform.device_count);
    }
    return &g_platform;
}

CLDevice CLGetDevice(CLPlatform platform) {
    (void)platform;
    int ordinal = clpt_state_device();
    if (ordinal < 0) {
        const char *env = getenv("CLPT_DEVICE");
        ordinal = env ? atoi(env) : 0;
    }
    if (ordinal < 0 || ordinal >= g_platform.device_count) {
        fprintf(stderr, "CLGetDevice: device %d requested, %d available\n", ordinal, g_platform.device_count);
        exit(EXIT_FAILURE);
    }
    if (verbose()) {
        for (int i = 0; i < g_platform.device_count; i++) {
            cudaDeviceProp p;
            HANDLE_ERR(cudaGetDeviceProperties(&p, i));
            printf("%s%d) %s (sm_%d%d, %d SMs)\n", i == ordinal ? "*" : " ", i, p.name, p.major, p.minor,
                   p.multiProcessorCount);
        }
    }
    return (CLDevice)(intptr_t)(ordinal + 1);
}

CLContext CLCreateContext(CLPlatform platform, CLDevice device) {
    (void)platform;
    const int ordinal = (int)(intptr_t)device - 1;
    HANDLE_ERR(cudaSetDevice(ordinal));
    HANDLE_ERR(cudaFree(0)); // force the primary context
    return device;
}

CLProgram CLBuildProgram(const char *filename, CLContext context, CLDevice device) {
    (void)filename;
    (void)context;
    (void)device;
    // The kernels are compiled for sm_100a only; on any other device this is
    // where the failure surfaces (cudaErrorNoKernelImageForDevice), the way a
    // clBuildProgram failure does in src/CLHandler.c:240-258.
    cudaFuncAttributes attr;
    HANDLE_ERR(cudaFuncGetAttributes(&attr, clpt_render_kernel_symbol()));
    if (verbose()) {
        printf("Program: render kernel, %d registers, binary sm_%d\n", attr.numRegs, attr.binaryVersion);
    }
    return (CLProgram)clpt_render_kernel_symbol();
}

CLQueue CLCreateQueue(CLContext context, CLDevice device) {
    (void)context;
    (void)device;
    cudaStream_t s;
    HANDLE_ERR(cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking));
    return (CLQueue)s;
}

CLKernel CLCreateKernel(const char *kernel_name, CLProgram program) {
    (void)program;
    if (kernel_name == NULL || strcmp(kernel_name, "render") != 0) {
        // CL_INVALID_KERNEL_NAME in the reference
        fprintf(stderr, "%s:%d: CUDA Error: no kernel named \"%s\" (the program has \"render\")\n", __FILE__,
                __LINE__, kernel_name ? kernel_name : "(null)");
        exit(EXIT_FAILURE);
    }
    return (CLKernel)clpt_render_kernel_symbol();
}

CLBuffer CLCreateBuffer(CLContext context, size_t size) {
    (void)context;
    void *p = NULL;
    HANDLE_ERR(cudaMalloc(&p, size));
    return p;
}

void CLReleaseBuffer(CLBuffer buffer) { HANDLE_ERR(cudaFree(buffer)); }

void CLWriteBuffer(CLQueue queue, CLBuffer dst, const void *src, size_t size) {
    HANDLE_ERR(cudaMemcpyAsync(dst, src, size, cudaMemcpyHostToDevice, (cudaStream_t)queue));
    HANDLE_ERR(cudaStreamSynchronize((cudaStream_t)queue));
}

void CLEnqueueKernel(unsigned int dim, size_t *global_size, size_t *local_size, CLQueue queue,
                     CLKernel kernel) {
    (void)local_size;
    (void)queue;
    if (dim != 2 || kernel != (CLKernel)clpt_render_kernel_symbol()) {
        fprintf(stderr, "CLEnqueueKernel: only the 2-D render kernel exists\n");
        exit(EXIT_FAILURE);
    }
    clpt_state_launch_frame((int)global_size[0], (int)global_size[1]);
}

} // extern "C"
