/* egl_present.c -- exercise the GL presentation hook (CLCreateImage(GLuint)) without a display.
 *
 * The reference renders into a GL texture shared with OpenCL (src/CLState.c:47-63,204-219,
 * src/GLHandler.c:164-211).  This program stands in for GLHandler.c on a headless GPU box: it
 * creates a surfaceless OpenGL context on the NVIDIA device through EGL, makes an RGBA8 texture the
 * size of the frame, hands its name to CLCreateImage, renders one frame with CLExecute (which maps
 * the texture, writes the frame as UNORM8 texels and unmaps it), reads the texture back with
 * glGetTexImage and compares it with CLReadImageRGBA8 -- the same frame, the same quantisation.
 *
 * The GPU boxes ship NVIDIA's EGL VENDOR library (libEGL_nvidia.so.0) but no libglvnd (no
 * libEGL.so.1 to load it through), and a vendor library exports a single entry point, __egl_Main,
 * meant to be called by libglvnd.  So this file plays libglvnd's part: it hands the vendor library
 * the small table of callbacks it expects (current context bookkeeping, error slot) and takes the
 * EGL and GL entry points from the table it gets back.  Struct layouts follow libglvnd's
 * src/EGL/libeglabi.h (ABI 0.1), declared here by hand -- no EGL or GL headers exist on these
 * machines.  Exit codes: 0 = presented and verified, 77 = no usable EGL/GL here (reason printed;
 * the test skips), anything else = a real failure.
 */
#include <dlfcn.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "CLState.h"
#include "clpt_host.h"

typedef void *EGLDisplay, *EGLContext, *EGLSurface, *EGLConfig, *EGLDeviceEXT;
typedef unsigned int EGLBoolean, EGLenum;
typedef int32_t EGLint;
typedef intptr_t EGLAttrib;
#define EGL_TRUE 1
#define EGL_FALSE 0
#define EGL_SUCCESS 0x3000
#define EGL_NONE 0x3038
#define EGL_OPENGL_API 0x30A2
#define EGL_OPENGL_ES_API 0x30A0
#define EGL_PLATFORM_DEVICE_EXT 0x313F
#define EGL_SURFACE_TYPE 0x3033
#define EGL_PBUFFER_BIT 0x0001
#define EGL_RENDERABLE_TYPE 0x3040
#define EGL_OPENGL_BIT 0x0008
#define EGL_CONTEXT_MAJOR_VERSION 0x3098
#define EGL_CONTEXT_MINOR_VERSION 0x30FB
#define EGL_DRAW 0x3059
#define EGL_READ 0x305A
#define EGL_EXTENSIONS 0x3055
#define EGL_VENDOR 0x3053

#define GL_TEXTURE_2D 0x0DE1
#define GL_RGBA 0x1908
#define GL_RGBA8 0x8058
#define GL_UNSIGNED_BYTE 0x1401
#define GL_TEXTURE_MIN_FILTER 0x2801
#define GL_TEXTURE_MAG_FILTER 0x2800
#define GL_NEAREST 0x2600
#define GL_NO_ERROR 0
#define GL_VENDOR 0x1F00
#define GL_RENDERER 0x1F01
#define GL_VERSION 0x1F02

typedef struct VendorInfo { int unused; } VendorInfo;
typedef void (*proc_t)(void);

/* libglvnd -> vendor */
typedef struct ApiExports {
    void (*threadInit)(void);
    EGLenum (*getCurrentApi)(void);
    VendorInfo *(*getCurrentVendor)(void);
    EGLContext (*getCurrentContext)(void);
    EGLDisplay (*getCurrentDisplay)(void);
    EGLSurface (*getCurrentSurface)(EGLint readDraw);
    proc_t (*fetchDispatchEntry)(VendorInfo *vendor, int index);
    void (*setEGLError)(EGLint errorCode);
    EGLBoolean (*setLastVendor)(VendorInfo *vendor);
    VendorInfo *(*getVendorFromDisplay)(EGLDisplay dpy);
    VendorInfo *(*getVendorFromDevice)(EGLDeviceEXT dev);
    void (*setVendorForDevice)(EGLDeviceEXT dev, VendorInfo *vendor);
    void *reserved[8];
} ApiExports;

/* vendor -> libglvnd */
typedef struct ApiImports {
    EGLDisplay (*getPlatformDisplay)(EGLenum platform, void *nativeDisplay, const EGLAttrib *attrib_list);
    EGLBoolean (*getSupportsAPI)(EGLenum api);
    const char *(*getVendorString)(int name);
    void *(*getProcAddress)(const char *procName);
    void *(*getDispatchAddress)(const char *procName);
    void (*setDispatchIndex)(const char *procName, int index);
    EGLBoolean (*isPatchSupported)(int type, int stubSize);
    EGLBoolean (*initiatePatch)(int type, int stubSize, void *lookupStubOffset);
    void (*releasePatch)(void);
    void (*patchThreadAttach)(void);
    EGLenum (*findNativeDisplayPlatform)(void *nativeDisplay);
    void *reserved[16];
} ApiImports;

static VendorInfo g_vendor;
static EGLenum g_api = EGL_OPENGL_ES_API;
static EGLContext g_ctx;
static EGLDisplay g_dpy;
static EGLSurface g_draw, g_read;
static EGLint g_err = EGL_SUCCESS;

static void x_threadInit(void) {}
static EGLenum x_getCurrentApi(void) { return g_api; }
static VendorInfo *x_getCurrentVendor(void) { return g_ctx ? &g_vendor : NULL; }
static EGLContext x_getCurrentContext(void) { return g_ctx; }
static EGLDisplay x_getCurrentDisplay(void) { return g_dpy; }
static EGLSurface x_getCurrentSurface(EGLint rd) { return rd == EGL_READ ? g_read : g_draw; }
static proc_t x_fetchDispatchEntry(VendorInfo *v, int i) { (void)v; (void)i; return NULL; }
static void x_setEGLError(EGLint e) { g_err = e; }
static EGLBoolean x_setLastVendor(VendorInfo *v) { (void)v; return EGL_TRUE; }
static VendorInfo *x_getVendorFromDisplay(EGLDisplay d) { (void)d; return &g_vendor; }
static VendorInfo *x_getVendorFromDevice(EGLDeviceEXT d) { (void)d; return &g_vendor; }
static void x_setVendorForDevice(EGLDeviceEXT d, VendorInfo *v) { (void)d; (void)v; }

static int skip(const char *why) {
    printf("SKIP: %s (last EGL error 0x%x)\n", why, (unsigned)g_err);
    return 77;
}

int main(int argc, char **argv) {
    setvbuf(stdout, NULL, _IONBF, 0);
    const int w = argc > 1 ? atoi(argv[1]) : 320, h = argc > 2 ? atoi(argv[2]) : 240;
    const char *names[] = { "libEGL_nvidia.so.0", "/usr/local/nvidia/lib64/libEGL_nvidia.so.0",
                            "/usr/lib/x86_64-linux-gnu/libEGL_nvidia.so.0" };
    void *lib = NULL;
    for (int i = 0; i < 3 && !lib; i++) lib = dlopen(names[i], RTLD_NOW | RTLD_GLOBAL);
    if (!lib) {
        printf("dlopen: %s\n", dlerror());
        return skip("no libEGL_nvidia.so.0");
    }
    typedef EGLBoolean (*egl_main_t)(uint32_t, const ApiExports *, VendorInfo *, ApiImports *);
    egl_main_t egl_main = (egl_main_t)dlsym(lib, "__egl_Main");
    if (!egl_main) return skip("no __egl_Main in the vendor library");
    ApiExports ex;
    memset(&ex, 0, sizeof ex);
    ex.threadInit = x_threadInit;
    ex.getCurrentApi = x_getCurrentApi;
    ex.getCurrentVendor = x_getCurrentVendor;
    ex.getCurrentContext = x_getCurrentContext;
    ex.getCurrentDisplay = x_getCurrentDisplay;
    ex.getCurrentSurface = x_getCurrentSurface;
    ex.fetchDispatchEntry = x_fetchDispatchEntry;
    ex.setEGLError = x_setEGLError;
    ex.setLastVendor = x_setLastVendor;
    ex.getVendorFromDisplay = x_getVendorFromDisplay;
    ex.getVendorFromDevice = x_getVendorFromDevice;
    ex.setVendorForDevice = x_setVendorForDevice;
    ApiImports im;
    memset(&im, 0, sizeof im);
    EGLBoolean ok = EGL_FALSE;
    const uint32_t versions[] = { (0u << 16) | 1u, (0u << 16) | 0u, (0u << 16) | 2u, (1u << 16) | 0u };
    uint32_t used = 0;
    for (int i = 0; i < 4 && !ok; i++) {
        memset(&im, 0, sizeof im);
        ok = egl_main(versions[i], &ex, &g_vendor, &im);
        used = versions[i];
    }
    if (!ok || !im.getProcAddress || !im.getPlatformDisplay) return skip("__egl_Main refused every vendor ABI version tried");
    printf("vendor ABI %u.%u accepted; vendor string: %s\n", used >> 16, used & 0xffff,
           im.getVendorString ? im.getVendorString(0) : "?");

#define EGLFN(ret, name, args) ret(*name) args = (ret(*) args)im.getProcAddress(#name); \
    if (!name) { printf("missing %s\n", #name); return skip("EGL entry point missing"); }
    EGLFN(EGLBoolean, eglQueryDevicesEXT, (EGLint, EGLDeviceEXT *, EGLint *))
    EGLFN(EGLBoolean, eglInitialize, (EGLDisplay, EGLint *, EGLint *))
    EGLFN(EGLBoolean, eglBindAPI, (EGLenum))
    EGLFN(EGLBoolean, eglChooseConfig, (EGLDisplay, const EGLint *, EGLConfig *, EGLint, EGLint *))
    EGLFN(EGLContext, eglCreateContext, (EGLDisplay, EGLConfig, EGLContext, const EGLint *))
    EGLFN(EGLBoolean, eglMakeCurrent, (EGLDisplay, EGLSurface, EGLSurface, EGLContext))
    EGLFN(const char *, eglQueryString, (EGLDisplay, EGLint))
    EGLFN(EGLBoolean, eglDestroyContext, (EGLDisplay, EGLContext))
    EGLFN(EGLBoolean, eglTerminate, (EGLDisplay))

    EGLDeviceEXT devs[16];
    EGLint ndev = 0;
    EGLDisplay dpy = NULL;
    EGLint major = 0, minor = 0;
    const EGLBoolean q = eglQueryDevicesEXT(16, devs, &ndev);
    printf("eglQueryDevicesEXT: returned %u, %d device(s), EGL error 0x%x\n", q, ndev, (unsigned)g_err);
    for (int d = 0; q && d < ndev && !dpy; d++) {
        EGLDisplay cand = im.getPlatformDisplay(EGL_PLATFORM_DEVICE_EXT, devs[d], NULL);
        if (cand && eglInitialize(cand, &major, &minor)) dpy = cand;
    }
    if (!dpy) { /* no device enumerated (a compute-only container?): the surfaceless platform, then the default display */
        const EGLenum platforms[] = { 0x31DD /* EGL_PLATFORM_SURFACELESS_MESA */, EGL_PLATFORM_DEVICE_EXT };
        for (int k = 0; k < 2 && !dpy; k++) {
            EGLDisplay cand = im.getPlatformDisplay(platforms[k], NULL, NULL);
            printf("getPlatformDisplay(0x%x, default): %p, EGL error 0x%x\n", platforms[k], cand, (unsigned)g_err);
            if (cand && eglInitialize(cand, &major, &minor)) dpy = cand;
            else if (cand) printf("  eglInitialize failed, EGL error 0x%x\n", (unsigned)g_err);
        }
    }
    if (!dpy) return skip("no EGL display could be initialised (no device enumerated: the container exposes the GPU for compute only)");
    printf("EGL %d.%d, vendor %s\n", major, minor, eglQueryString(dpy, EGL_VENDOR));
    if (!eglBindAPI(EGL_OPENGL_API)) return skip("eglBindAPI(EGL_OPENGL_API) failed");
    g_api = EGL_OPENGL_API;
    const EGLint cfg_attr[] = { EGL_SURFACE_TYPE, EGL_PBUFFER_BIT, EGL_RENDERABLE_TYPE, EGL_OPENGL_BIT, EGL_NONE };
    EGLConfig cfg = NULL;
    EGLint ncfg = 0;
    if (!eglChooseConfig(dpy, cfg_attr, &cfg, 1, &ncfg) || ncfg < 1) return skip("no EGL config for OpenGL");
    const EGLint ctx_attr[] = { EGL_CONTEXT_MAJOR_VERSION, 4, EGL_CONTEXT_MINOR_VERSION, 5, EGL_NONE };
    EGLContext ctx = eglCreateContext(dpy, cfg, NULL, ctx_attr);
    if (!ctx) return skip("eglCreateContext(OpenGL 4.5) failed");
    if (!eglMakeCurrent(dpy, NULL, NULL, ctx)) return skip("eglMakeCurrent (surfaceless) failed");
    g_ctx = ctx;
    g_dpy = dpy;
    g_draw = g_read = NULL;

#define GLFN(ret, name, args) ret(*name) args = (ret(*) args)im.getProcAddress(#name); \
    if (!name) { printf("missing %s\n", #name); return skip("GL entry point missing"); }
    GLFN(const unsigned char *, glGetString, (unsigned))
    GLFN(void, glGenTextures, (int, unsigned *))
    GLFN(void, glBindTexture, (unsigned, unsigned))
    GLFN(void, glTexImage2D, (unsigned, int, int, int, int, int, unsigned, unsigned, const void *))
    GLFN(void, glTexParameteri, (unsigned, unsigned, int))
    GLFN(void, glGetTexImage, (unsigned, int, unsigned, unsigned, void *))
    GLFN(void, glFinish, (void))
    GLFN(unsigned, glGetError, (void))
    GLFN(void, glDeleteTextures, (int, const unsigned *))
    const unsigned char *glver = glGetString(GL_VERSION);
    if (!glver) return skip("glGetString returned NULL: the context is not current for GL calls made this way");
    printf("GL %s / %s\n", (const char *)glver, (const char *)glGetString(GL_RENDERER));
    printf("CONTEXT: an OpenGL context is current\n");

    /* the texture GLHandler.c makes: RGBA8, GL_TEXTURE_2D, level 0 (src/GLHandler.c:177-185) */
    unsigned tex = 0;
    glGenTextures(1, &tex);
    glBindTexture(GL_TEXTURE_2D, tex);
    unsigned char *zeros = calloc((size_t)w * h, 4);
    glTexImage2D(GL_TEXTURE_2D, 0, GL_RGBA8, w, h, 0, GL_RGBA, GL_UNSIGNED_BYTE, zeros);
    glTexParameteri(GL_TEXTURE_2D, GL_TEXTURE_MIN_FILTER, GL_NEAREST);
    glTexParameteri(GL_TEXTURE_2D, GL_TEXTURE_MAG_FILTER, GL_NEAREST);
    glFinish();
    if (glGetError() != GL_NO_ERROR) return skip("texture creation raised a GL error");

    /* from here on: the reference's call order, with the GL texture as the render target */
    kd *models = new_list(sizeof(kd));
    kd model;
    if (argc > 3) {
        if (LoadModel(argv[3], &model)) return 2;
    } else {
        printf("usage: egl_present W H model.obj|.kd\n");
        return 2;
    }
    vector_append(models, model);
    CLInit("src/kernel.cl", "render");
    CLCreateImage(tex); /* cudaGraphicsGLRegisterImage on the current context */
    CLSetMeshes(models);
    CLSetRenderParams(CLPT_MODE_MIRROR, 2, 1, 0, 0);
    Camera cam = { 0.1f, 1.0f, 1.0471976f, Vector3(0, 0.9f, -1.7f), Vector3(0, -0.42f, 0.9075f) };
    Matrix m = cam_matrix(cam, h);
    CLSetCameraMatrix(m);
    CLExecute(w, h); /* map, write UNORM8 texels, unmap (src/CLState.c:207-218) */
    unsigned char *from_gl = malloc((size_t)w * h * 4), *from_cl = malloc((size_t)w * h * 4);
    glBindTexture(GL_TEXTURE_2D, tex);
    glGetTexImage(GL_TEXTURE_2D, 0, GL_RGBA, GL_UNSIGNED_BYTE, from_gl);
    glFinish();
    CLReadImageRGBA8(from_cl, (size_t)w * h * 4);
    size_t differ = 0, nonwhite = 0;
    for (size_t i = 0; i < (size_t)w * h * 4; i++) {
        differ += from_gl[i] != from_cl[i];
        nonwhite += from_cl[i] != 255;
    }
    printf("texture vs CLReadImageRGBA8: %zu of %zu bytes differ; %zu bytes are not 255 (the frame shows the scene)\n",
           differ, (size_t)w * h * 4, nonwhite);
    CLDeleteImage();
    CLTerminate();
    glDeleteTextures(1, &tex);
    eglMakeCurrent(dpy, NULL, NULL, NULL);
    g_ctx = NULL;
    eglDestroyContext(dpy, ctx);
    eglTerminate(dpy);
    if (differ != 0 || nonwhite == 0) {
        printf("FAIL\n");
        return 1;
    }
    printf("PRESENTED: the GL texture holds the frame\n");
    return 0;
}
