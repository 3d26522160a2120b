/* CLHandler.h -- thin runtime wrapper under CLState, CUDA-shaped.
 *
 * The reference's include/CLHandler.h:6-25 wraps eight OpenCL runtime calls
 * and returns cl_* handles.  Those handle types do not exist here (no OpenCL
 * anywhere in this library); the same eight steps are exposed with opaque
 * handles over the CUDA runtime.  They are internal plumbing for CLState --
 * the application never calls them (in the reference only src/CLState.c does).
 *
 * Errors follow include/error.h:3: HANDLE_ERR(e) prints
 * "file:line: CUDA Error: NAME" and exits.
 */
#ifndef CLHANDLER_H
#define CLHANDLER_H

#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct CLPlatform_ *CLPlatform; /* the CUDA driver/runtime pair */
typedef void *CLDevice;                 /* device ordinal + 1 */
typedef void *CLContext;                /* primary context of the device */
typedef void *CLProgram;                /* the kernels compiled into libclpt.so */
typedef void *CLQueue;                  /* cudaStream_t */
typedef void *CLKernel;                 /* a named entry of the program */
typedef void *CLBuffer;                 /* device pointer */

/* src/CLHandler.c:56-108 prompted on stdin when several platforms/devices
 * exist; here the device is chosen by ordinal (CLSelectDevice/$CLPT_DEVICE)
 * and the list is only printed when $CLPT_VERBOSE is set. */
CLPlatform CLGetPlatform(void);                                     /* CLHandler.h:6-7 */
CLDevice CLGetDevice(CLPlatform platform);                          /* :8-9 */
CLContext CLCreateContext(CLPlatform platform, CLDevice device);    /* :10-11 */
/* Checks that the sm_100a image of the built-in program loads on `device`
 * (the moral equivalent of clBuildProgram failing with a build log);
 * `filename` is ignored. */
CLProgram CLBuildProgram(const char *filename, CLContext context, CLDevice device); /* :12-13 */
CLQueue CLCreateQueue(CLContext context, CLDevice device);          /* :14-15 */
CLKernel CLCreateKernel(const char *kernel_name, CLProgram program); /* :16-17; "render" */
CLBuffer CLCreateBuffer(CLContext context, size_t size);            /* :18-19 */
void CLReleaseBuffer(CLBuffer buffer);
void CLWriteBuffer(CLQueue queue, CLBuffer dst, const void *src, size_t size); /* blocking */
/* Launch `kernel` over a dim-dimensional global range; local_size may be NULL
 * (the reference always passes NULL, src/CLState.c:209-211). */
void CLEnqueueKernel(unsigned int dim, size_t *global_size, size_t *local_size,
                     CLQueue queue, CLKernel kernel);               /* :20-25 */

/* include/error.h:5-9 */
#define HANDLE_ERR(err) handle_err((err), __FILE__, __LINE__)
const char *err_string(int error);
void handle_err(int err, const char *file, int line);

#ifdef __cplusplus
}
#endif

#endif /* CLHANDLER_H */
