/* model_io.c -- OBJ / .kd model loading and the Euler integrator
 * (clpt_host.h "models", "physics").
 *
 * LoadModel mirrors the reference's dispatch on the file extension
 * (src/model.c:147-176): ".obj" is parsed, a kd-tree is built and cached as
 * "<stem>.kd"; ".kd" is read back directly.  The reference parses OBJ text
 * with the vendored tinyobj_loader_c.h; this reader is a small line scanner
 * for the subset synthetic scenes use (v, vn, vt, f) that yields the SAME
 * lists the reference's loader glue builds (src/model.c:109-133):
 *   verts : one Vector3 per "v" line, file order
 *   norms : one Vector3 per "vn" line
 *   tris  : one cl_int3 {v, vn, vt} per triangle corner, polygons fanned as
 *           (c0, c[k-1], c[k]); 1-based -> 0-based; negative = relative to
 *           the elements seen so far; an absent index stays "invalid"
 *           (INT_MIN + count, i.e. negative -- the kernel tests vn >= 0).
 */
#include <ctype.h>
#include <limits.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "clpt_host.h"

static int g_build_depth = KD_REF_DEPTH;
static int g_build_nbins = KD_REF_NBINS;

void
kd_set_build_params(int depth, int nbins) {
    g_build_depth = depth;
    g_build_nbins = nbins;
}

/* Decimal text -> double.  Deliberately NOT strtod: to load the same floats
 * as the reference's OBJ path, the digits are combined the way its loader
 * does (tinyobj_loader_c.h:296-429): integer digits by *10, fraction digit k
 * scaled by 0.1 multiplied k times, a decimal exponent applied as 5^e * 2^e.
 * The result is narrowed to float by the caller. */
static double
scan_real(const char **cursor) {
    const char *p = *cursor;
    while (*p == ' ' || *p == '\t') {
        p++;
    }
    const char *tok = p;
    while (*p && *p != ' ' && *p != '\t' && *p != '\r' && *p != '\n') {
        p++;
    }
    const char *end = p;
    *cursor = end;

    p = tok;
    if (p >= end) {
        return 0.0;
    }
    int neg = 0;
    if (*p == '+' || *p == '-') {
        neg = *p == '-';
        p++;
    } else if (!isdigit((unsigned char)*p)) {
        return 0.0;
    }
    double mant = 0.0;
    int ndig = 0;
    while (p < end && isdigit((unsigned char)*p)) {
        mant = mant * 10 + (*p - '0');
        p++;
        ndig++;
    }
    if (ndig == 0) {
        return 0.0;
    }
    int expo = 0, expo_neg = 0;
    if (p < end && *p == '.') {
        p++;
        int place = 1;
        while (p < end && isdigit((unsigned char)*p)) {
            double scale = 1.0;
            for (int k = 0; k < place; k++) {
                scale *= 0.1;
            }
            mant += (*p - '0') * scale;
            place++;
            p++;
        }
    }
    if (p < end && (*p == 'e' || *p == 'E')) {
        p++;
        if (p < end && (*p == '+' || *p == '-')) {
            expo_neg = *p == '-';
            p++;
        } else if (!(p < end && isdigit((unsigned char)*p))) {
            return 0.0;
        }
        int nd = 0;
        while (p < end && isdigit((unsigned char)*p)) {
            expo = expo * 10 + (*p - '0');
            p++;
            nd++;
        }
        if (nd == 0) {
            return 0.0;
        }
    }
    double five = 1.0, two = 1.0;
    for (int k = 0; k < expo; k++) {
        five *= 5.0;
        two *= 2.0;
    }
    if (expo_neg) {
        five = 1.0 / five;
        two = 1.0 / two;
    }
    return (neg ? -1 : 1) * (mant * five * two);
}

static int
scan_int(const char **cursor) {
    const char *p = *cursor;
    int neg = 0, val = 0;
    if (*p == '+' || *p == '-') {
        neg = *p == '-';
        p++;
    }
    while (isdigit((unsigned char)*p)) {
        val = val * 10 + (*p - '0');
        p++;
    }
    *cursor = p;
    return neg ? -val : val;
}

static int
resolve_index(int raw, size_t seen) {
    if (raw > 0) {
        return raw - 1;
    }
    if (raw == 0) {
        return 0;
    }
    return (int)seen + raw; /* relative; also maps INT_MIN -> a negative */
}

static void
skip_to_sep(const char **cursor) {
    const char *p = *cursor;
    while (*p && *p != '/' && *p != ' ' && *p != '\t' && *p != '\r' && *p != '\n') {
        p++;
    }
    *cursor = p;
}

/* one face corner: i, i/j, i//k, i/j/k */
static void
scan_corner(const char **cursor, int raw[3]) {
    raw[0] = raw[1] = raw[2] = INT_MIN; /* v, vn, vt */
    raw[0] = scan_int(cursor);
    skip_to_sep(cursor);
    if (**cursor != '/') {
        return;
    }
    (*cursor)++;
    if (**cursor == '/') {
        (*cursor)++;
        raw[1] = scan_int(cursor);
        skip_to_sep(cursor);
        return;
    }
    raw[2] = scan_int(cursor);
    skip_to_sep(cursor);
    if (**cursor != '/') {
        return;
    }
    (*cursor)++;
    raw[1] = scan_int(cursor);
    skip_to_sep(cursor);
}

int
load_obj_lists(const char *filename, Vector3 **verts_out, Vector3 **norms_out,
               cl_int3 **tris_out) {
    FILE *f = fopen(filename, "rb");
    if (f == NULL) {
        perror(filename);
        return 1;
    }
    fseek(f, 0, SEEK_END);
    long len = ftell(f);
    fseek(f, 0, SEEK_SET);
    char *text = malloc((size_t)len + 1);
    if (text == NULL) {
        perror("malloc");
        exit(EXIT_FAILURE);
    }
    if (fread(text, 1, (size_t)len, f) != (size_t)len) {
        fprintf(stderr, "%s: error reading from file\n", filename);
        fclose(f);
        free(text);
        return 1;
    }
    text[len] = '\0';
    fclose(f);

    Vector3 *verts = new_list(0), *norms = new_list(0);
    cl_int3 *tris = new_list(0);
    size_t nv = 0, nn = 0, nt = 0;
    const char *line = text;
    while (*line) {
        const char *eol = line;
        while (*eol && *eol != '\n') {
            eol++;
        }
        const char *p = line;
        while (*p == ' ' || *p == '\t') {
            p++;
        }
        if (p[0] == 'v' && (p[1] == ' ' || p[1] == '\t')) {
            p += 2;
            float x = (float)scan_real(&p), y = (float)scan_real(&p),
                  z = (float)scan_real(&p);
            vector_append(verts, Vector3(x, y, z));
            nv++;
        } else if (p[0] == 'v' && p[1] == 'n' && (p[2] == ' ' || p[2] == '\t')) {
            p += 3;
            float x = (float)scan_real(&p), y = (float)scan_real(&p),
                  z = (float)scan_real(&p);
            vector_append(norms, Vector3(x, y, z));
            nn++;
        } else if (p[0] == 'v' && p[1] == 't' && (p[2] == ' ' || p[2] == '\t')) {
            nt++;
        } else if (p[0] == 'f' && (p[1] == ' ' || p[1] == '\t')) {
            p += 2;
            cl_int3 first = { { 0, 0, 0, 0 } }, prev = first;
            int k = 0;
            for (;;) {
                while (*p == ' ' || *p == '\t') {
                    p++;
                }
                if (p >= eol || *p == '\r' || *p == '\n' || *p == '\0') {
                    break;
                }
                int raw[3];
                scan_corner(&p, raw);
                cl_int3 c = { { resolve_index(raw[0], nv), resolve_index(raw[1], nn),
                                resolve_index(raw[2], nt), 0 } };
                if (k == 0) {
                    first = c;
                } else if (k >= 2) {
                    vector_append(tris, first);
                    vector_append(tris, prev);
                    vector_append(tris, c);
                }
                prev = c;
                k++;
            }
        }
        line = *eol ? eol + 1 : eol;
    }
    free(text);
    *verts_out = verts;
    *norms_out = norms;
    *tris_out = tris;
    return 0;
}

int
write_obj(const char *filename, const Vector3 *verts, const Vector3 *norms,
          const cl_int3 *tris) {
    FILE *f = fopen(filename, "wb");
    if (f == NULL) {
        perror(filename);
        return 1;
    }
    size_t nv = vector_length(verts), nn = norms ? vector_length(norms) : 0;
    size_t nc = vector_length(tris);
    for (size_t i = 0; i < nv; i++) {
        fprintf(f, "v %.6f %.6f %.6f\n", verts[i].s[0], verts[i].s[1], verts[i].s[2]);
    }
    for (size_t i = 0; i < nn; i++) {
        fprintf(f, "vn %.6f %.6f %.6f\n", norms[i].s[0], norms[i].s[1], norms[i].s[2]);
    }
    for (size_t i = 0; i + 2 < nc; i += 3) {
        fputc('f', f);
        for (int c = 0; c < 3; c++) {
            int v = tris[i + c].s[0], vn = tris[i + c].s[1];
            if (vn >= 0) {
                fprintf(f, " %d//%d", v + 1, vn + 1);
            } else {
                fprintf(f, " %d", v + 1);
            }
        }
        fputc('\n', f);
    }
    return fclose(f) != 0;
}

int
LoadModel(const char *filename, kd *model) {
    const char *ext = strrchr(filename, '.');
    if (ext != NULL && strcmp(ext, ".obj") == 0) {
        Vector3 *verts, *norms;
        cl_int3 *tris;
        if (load_obj_lists(filename, &verts, &norms, &tris)) {
            return 1;
        }
        size_t stem_len = (size_t)(ext - filename);
        char *stem = calloc(stem_len + 1, 1);
        memcpy(stem, filename, stem_len);
        *model = build_kd_ex(tris, verts, norms, stem, g_build_depth, g_build_nbins);
        free(stem);
        return 0;
    }
    if (ext != NULL && strcmp(ext, ".kd") == 0) {
        return parse_kd(filename, model);
    }
    fprintf(stderr, "Unrecognized filetype: \"%s\"\n", filename);
    fprintf(stderr, "Supported filetypes are: \".obj\", \".kd\"\n");
    return 1;
}

/* ---- physics (src/physics.c:24-64): explicit Euler over pointer pairs ---- */

typedef struct phys_pair {
    Vector3 *pos, *vel;
} phys_pair;

static phys_pair *g_pairs = NULL;

void
AddPhysObject(Vector3 *position, Vector3 *velocity) {
    if (g_pairs == NULL) {
        g_pairs = new_list(sizeof(*g_pairs));
    }
    phys_pair p = { position, velocity };
    vector_append(g_pairs, p);
}

void
PhysStep(double stepSize) {
    size_t n = g_pairs ? vector_length(g_pairs) : 0;
    for (size_t i = 0; i < n; i++) {
        /* the step narrows to float when passed to vec_scaled (physics.c:51-52) */
        *g_pairs[i].pos = vec_add(*g_pairs[i].pos, vec_scaled(*g_pairs[i].vel, (vec_t)stepSize));
    }
}

void
PhysTerminate(void) {
    delete_list(g_pairs);
    g_pairs = NULL;
}
