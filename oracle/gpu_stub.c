/*
 * TEST INFRASTRUCTURE ONLY.  Never linked into the product.
 *
 * Stub definitions for the GPU C-ABI symbols the reference's aligner.cpp
 * refers to (declared in the reference's aligner_gpu.h:32-117).  The reference
 * CPU path (`ghostm aln` without -D) never calls them, so they abort loudly if
 * they are ever reached.  Linking them lets oracle/Makefile build the
 * reference's CPU aligner from the sources under /root/reference without nvcc
 * and without the reference's own (unbuildable here) aligner_gpu.cu.
 */
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

static void die(const char *name) {
  fprintf(stderr, "oracle/gpu_stub.c: %s called - the CPU oracle must run without -D\n", name);
  abort();
}

int InitGpu(void) { die("InitGpu"); return 1; }
int CheckGpuMemory(uint32_t a, uint32_t b, uint32_t c, uint32_t d, uint32_t e, uint32_t f) {
  (void)a; (void)b; (void)c; (void)d; (void)e; (void)f; die("CheckGpuMemory"); return 1; }
int SetOptionGpu(uint32_t a, int m[], int d) { (void)a; (void)m; (void)d; die("SetOptionGpu"); return 1; }
void printGpuInfo(int d) { (void)d; die("printGpuInfo"); }
int SetQueryGpu(uint8_t s[], uint32_t n, uint32_t l) { (void)s; (void)n; (void)l; die("SetQueryGpu"); return 1; }
int SetDbGpu(uint8_t s[], uint32_t sl, uint32_t kc[], uint32_t kl, uint32_t p[], uint32_t pl) {
  (void)s; (void)sl; (void)kc; (void)kl; (void)p; (void)pl; die("SetDbGpu"); return 1; }
uint32_t SearchNextGpu(uint32_t a, uint32_t b, uint32_t c, uint32_t d, uint32_t e, uint32_t f,
                       uint32_t g, uint32_t h, uint32_t *i, uint32_t *j) {
  (void)a; (void)b; (void)c; (void)d; (void)e; (void)f; (void)g; (void)h; (void)i; (void)j;
  die("SearchNextGpu"); return 0; }
void CalculateScoreGpu(uint32_t a, uint32_t b, uint32_t c, uint32_t s[], uint32_t e[],
                       uint32_t f, uint32_t g, int h, int i) {
  (void)a; (void)b; (void)c; (void)s; (void)e; (void)f; (void)g; (void)h; (void)i;
  die("CalculateScoreGpu"); }
int FreeGpu(void) { die("FreeGpu"); return 1; }
