// gl_interop.cu -- the presentation hook: CLCreateImage(GLuint texture).
//
// The reference renders into a GL texture shared with OpenCL
// (clCreateFromGLTexture, src/CLState.c:47-58; acquire/release around the launch,
// :60-63,204-219).  The CUDA equivalent: register the texture once
// (cudaGraphicsGLRegisterImage), and per frame map it, wrap its array in a surface
// object, convert the float4 frame to the texture's RGBA8 texels, unmap.
//
// No OpenGL header exists on the build machines, so the one interop entry point
// and the one GL enum used are declared here by hand (values from the Khronos
// registry / cuda_gl_interop.h); the function itself lives in the CUDA runtime
// that is linked statically.  Without a current GL context the registration
// fails and the library aborts with the CUDA error, like any other failure.
// There is no display on the build or bench machines, so this path is compiled
// and linked but not exercised by the tests.
#include <cuda_runtime.h>

#include "CLHandler.h"
#include "clpt_device.cuh"

extern "C" cudaError_t cudaGraphicsGLRegisterImage(struct cudaGraphicsResource **resource, unsigned int image,
                                                   unsigned int target, unsigned int flags);
#define CLPT_GL_TEXTURE_2D 0x0DE1

namespace {

cudaGraphicsResource *g_resource = nullptr;

__global__ void present_kernel(cudaSurfaceObject_t surf, const float4 *__restrict__ frame, int width, int height) {
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y * blockDim.y + threadIdx.y;
    if (x >= width || y >= height) return;
    const float4 c = frame[(size_t)y * width + x];
    const uchar4 texel = make_uchar4((unsigned char)clpt_to_unorm8(c.x), (unsigned char)clpt_to_unorm8(c.y),
                                     (unsigned char)clpt_to_unorm8(c.z), (unsigned char)clpt_to_unorm8(c.w));
    surf2Dwrite(texel, surf, x * (int)sizeof(uchar4), y);
}

} // namespace

void clpt_gl_register(unsigned int texture) {
    if (g_resource) HANDLE_ERR(cudaGraphicsUnregisterResource(g_resource));
    g_resource = nullptr;
    HANDLE_ERR(cudaGraphicsGLRegisterImage(&g_resource, texture, CLPT_GL_TEXTURE_2D,
                                           cudaGraphicsRegisterFlagsSurfaceLoadStore |
                                               cudaGraphicsRegisterFlagsWriteDiscard));
}

void clpt_gl_unregister(void) {
    if (g_resource) HANDLE_ERR(cudaGraphicsUnregisterResource(g_resource));
    g_resource = nullptr;
}

bool clpt_gl_registered(void) { return g_resource != nullptr; }

// Size of the registered texture's level 0.
void clpt_gl_size(int *width, int *height, cudaStream_t stream) {
    HANDLE_ERR(cudaGraphicsMapResources(1, &g_resource, stream));
    cudaArray_t array;
    HANDLE_ERR(cudaGraphicsSubResourceGetMappedArray(&array, g_resource, 0, 0));
    cudaChannelFormatDesc desc;
    cudaExtent extent;
    unsigned int flags;
    HANDLE_ERR(cudaArrayGetInfo(&desc, &extent, &flags, array));
    *width = (int)extent.width;
    *height = (int)extent.height;
    HANDLE_ERR(cudaGraphicsUnmapResources(1, &g_resource, stream));
}

// acquire -> write -> release, the shape of src/CLState.c:207-218
void clpt_gl_present(const float4 *frame, int width, int height, cudaStream_t stream) {
    HANDLE_ERR(cudaGraphicsMapResources(1, &g_resource, stream));
    cudaArray_t array;
    HANDLE_ERR(cudaGraphicsSubResourceGetMappedArray(&array, g_resource, 0, 0));
    cudaResourceDesc rd = {};
    rd.resType = cudaResourceTypeArray;
    rd.res.array.array = array;
    cudaSurfaceObject_t surf = 0;
    HANDLE_ERR(cudaCreateSurfaceObject(&surf, &rd));
    dim3 block(32, 8), grid((width + 31) / 32, (height + 7) / 8);
    present_kernel<<<grid, block, 0, stream>>>(surf, frame, width, height);
    HANDLE_ERR(cudaGetLastError());
    HANDLE_ERR(cudaStreamSynchronize(stream));
    HANDLE_ERR(cudaDestroySurfaceObject(surf));
    HANDLE_ERR(cudaGraphicsUnmapResources(1, &g_resource, stream));
}
