/* headless_host.c -- a C host written against the reference's call order
 * (src/game.c:219-260): load model -> CLInit -> CLSetMeshes -> per frame
 * { CLSetCameraMatrix, CLSetObjects, CLExecute, PhysStep } -> CLTerminate.
 *
 *   gcc -std=c11 -Iinclude examples/headless_host.c -Lclpathtracer_b200 -lclpt \
 *       -Wl,-rpath,$PWD/clpathtracer_b200 -lm -o headless_host
 *   ./headless_host scene.obj 640 480 out.ppm
 */
#include <stdio.h>
#include <stdlib.h>

#include "CLState.h"
#include "clpt_host.h"

int
main(int argc, char **argv) {
    if (argc < 5) {
        fprintf(stderr, "usage: %s model.(obj|kd) width height out.ppm [frames]\n", argv[0]);
        return 2;
    }
    int w = atoi(argv[2]), h = atoi(argv[3]);
    int frames = argc > 5 ? atoi(argv[5]) : 3;

    kd *models = new_list(sizeof(kd));
    kd model;
    if (LoadModel(argv[1], &model)) {
        return 1;
    }
    vector_append(models, model);

    CLInit("src/kernel.cl", "render");
    CLCreateImageHeadless(w, h);
    CLSetMeshes(models);                              /* takes models[0]'s five lists */
    CLSetRenderParams(CLPT_MODE_MIRROR, 2, 1, 0, 0);  /* primary + 1 mirror bounce */

    Camera cam = { 0.1f, 1.0f, 1.0471976f, Vector3(0, 0.9f, -1.7f), Vector3(0, -0.42f, 0.9075f) };
    Vector3 vel = Vector3(0, 0, 0.05f);
    AddPhysObject(&cam.Position, &vel);               /* the camera is the only physics object, src/game.c:278 */
    for (int f = 0; f < frames; f++) {
        CLSetCameraMatrix(cam_matrix(cam, h));
        CLSetObjects(NULL, 0);
        CLExecute(w, h);
        PhysStep(0.016);
    }
    float *frame = malloc((size_t)w * h * 16);
    CLReadImage(frame, (size_t)w * h * 16);
    printf("frame %dx%d rendered in %.3f ms on %s\n", w, h, CLLastKernelMs(), CLDeviceName());

    FILE *out = fopen(argv[4], "wb");
    if (out) {
        fprintf(out, "P6\n%d %d\n255\n", w, h);
        for (int y = h - 1; y >= 0; y--) {            /* image row 0 is the bottom (y up) */
            for (int x = 0; x < w; x++) {
                const float *p = frame + 4 * ((size_t)y * w + x);
                for (int c = 0; c < 3; c++) {
                    float v = p[c] < 0 ? 0 : (p[c] > 1 ? 1 : p[c]);
                    fputc((int)(v * 255.0f + 0.5f), out);
                }
            }
        }
        fclose(out);
    }
    free(frame);
    PhysTerminate();
    CLTerminate();        /* frees the model's lists, src/CLState.c:221-225 */
    delete_list(models);  /* the caller frees only the outer list, src/game.c:177-178 */
    return 0;
}
