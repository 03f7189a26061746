// io.cpp -- reader / writer of the reference's text data file, main.cu:29-67:
//   "n_u N m num_iterations L" then M_G (n*m), g_P (n), G_L (n*m), p_D (m),
//   theta[num_iterations], beta[num_iterations]; whitespace-separated %f-parsable floats.
// The operator layout inside the file is whatever the consumer expects (the shipped kernels
// read the flipped one, kernel_functions.cu:50,180); this code does not reorder anything.
// Unlike readData() every fopen/fscanf result is checked (the reference ignores them, main.cu:32).
#include <cstdio>
#include <cstdlib>
#include <cstring>

#include "gpad.h"

namespace {

bool read_floats(FILE* fp, float* dst, size_t count) {
    for (size_t i = 0; i < count; ++i)
        if (fscanf(fp, "%f", dst + i) != 1) return false;
    return true;
}

bool write_floats(FILE* fp, const float* src, size_t count) {
    for (size_t i = 0; i < count; ++i)
        if (fprintf(fp, "%.9g\n", (double)src[i]) < 0) return false;
    return true;
}

}  // namespace

extern "C" {

void gpad_file_free(gpad_file_t* f) {
    if (!f) return;
    free(f->M_G); free(f->g_P); free(f->G_L); free(f->p_D); free(f->theta); free(f->beta);
    memset(f, 0, sizeof(*f));
}

int gpad_file_read(const char* path, gpad_file_t* out) {
    if (!path || !out) return GPAD_ERR_INVALID_ARG;
    memset(out, 0, sizeof(*out));
    FILE* fp = fopen(path, "r");
    if (!fp) return GPAD_ERR_IO;
    int rc = GPAD_ERR_IO;
    if (fscanf(fp, "%d %d %d %d %f", &out->n_u, &out->N, &out->m, &out->num_iterations, &out->L) == 5 &&
        out->n_u > 0 && out->N > 0 && out->m > 0 && out->num_iterations >= 0) {
        const size_t n = (size_t)out->n_u * out->N, m = (size_t)out->m, it = (size_t)out->num_iterations;
        out->M_G = (float*)malloc(sizeof(float) * n * m);
        out->g_P = (float*)malloc(sizeof(float) * n);
        out->G_L = (float*)malloc(sizeof(float) * n * m);
        out->p_D = (float*)malloc(sizeof(float) * m);
        out->theta = (float*)malloc(sizeof(float) * (it ? it : 1));
        out->beta = (float*)malloc(sizeof(float) * (it ? it : 1));
        if (out->M_G && out->g_P && out->G_L && out->p_D && out->theta && out->beta &&
            read_floats(fp, out->M_G, n * m) && read_floats(fp, out->g_P, n) && read_floats(fp, out->G_L, n * m) &&
            read_floats(fp, out->p_D, m) && read_floats(fp, out->theta, it) && read_floats(fp, out->beta, it))
            rc = GPAD_OK;
    }
    fclose(fp);
    if (rc != GPAD_OK) gpad_file_free(out);
    return rc;
}

int gpad_file_write(const char* path, const gpad_file_t* in) {
    if (!path || !in || !in->M_G || !in->g_P || !in->G_L || !in->p_D) return GPAD_ERR_INVALID_ARG;
    FILE* fp = fopen(path, "w");
    if (!fp) return GPAD_ERR_IO;
    const size_t n = (size_t)in->n_u * in->N, m = (size_t)in->m, it = (size_t)in->num_iterations;
    bool ok = fprintf(fp, "%d %d %d %d %.9g\n", in->n_u, in->N, in->m, in->num_iterations, (double)in->L) > 0;
    ok = ok && write_floats(fp, in->M_G, n * m) && write_floats(fp, in->g_P, n) && write_floats(fp, in->G_L, n * m) &&
         write_floats(fp, in->p_D, m) && write_floats(fp, in->theta, it) && write_floats(fp, in->beta, it);
    ok = (fclose(fp) == 0) && ok;
    return ok ? GPAD_OK : GPAD_ERR_IO;
}

}  // extern "C"
