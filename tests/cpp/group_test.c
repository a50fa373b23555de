/* group_test.c — a plain C caller of the multi-GPU part of the C ABI (include/b200sa.h, b200sa_group_*): one text sharded over
 * the listed devices, the reference's three calls with host buffers.  Prints digests that tests/test_c_group.py compares
 * with the oracle.
 *   group_test <input file> <device,device,...> */
#include <b200sa.h>

#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

static uint64_t fnv1a64(const void* data, size_t bytes)
{
    const uint8_t* p = (const uint8_t*)data;
    uint64_t h = 0xcbf29ce484222325ull;
    size_t i;
    for (i = 0; i < bytes; ++i) { h ^= p[i]; h *= 0x100000001b3ull; }
    return h;
}

int main(int argc, char** argv)
{
    int devices[16], count = 0, rc;
    b200sa_group* gpus = NULL;
    FILE* f;
    long n;
    uint8_t *text, *work;
    int32_t *sa, sentinel = 0;
    char* tok;
    if (argc < 3) { fprintf(stderr, "usage: group_test <input file> <device,device,...>\n"); return 2; }
    for (tok = strtok(argv[2], ","); tok && count < 16; tok = strtok(NULL, ",")) devices[count++] = atoi(tok);
    f = fopen(argv[1], "rb");
    if (!f) { fprintf(stderr, "cannot open %s\n", argv[1]); return 2; }
    fseek(f, 0, SEEK_END); n = ftell(f); fseek(f, 0, SEEK_SET);
    text = (uint8_t*)malloc((size_t)n + 1); work = (uint8_t*)malloc((size_t)n + 1); sa = (int32_t*)malloc(((size_t)n + 1) * 4);
    if (fread(text, 1, (size_t)n, f) != (size_t)n) return 2;
    fclose(f);
    rc = b200sa_group_create(&gpus, devices, count);
    if (rc) { printf("ERROR %d %s\n", rc, b200sa_last_error()); return 1; }
    printf("GROUP %d\n", b200sa_group_size(gpus));
    rc = b200sa_group_suffix_array(gpus, text, n, sa);                       /* make_suffix_array */
    if (rc) { printf("ERROR %d %s\n", rc, b200sa_last_error()); return 1; }
    printf("SA %016llx %ld\n", (unsigned long long)fnv1a64(sa, ((size_t)n + 1) * 4), n + 1);
    memcpy(work, text, (size_t)n);
    rc = b200sa_group_bwt(gpus, work, n, &sentinel);                         /* forward_burrows_wheeler_transform, in place */
    if (rc) { printf("ERROR %d %s\n", rc, b200sa_last_error()); return 1; }
    printf("BWT %016llx %d\n", (unsigned long long)fnv1a64(work, (size_t)n), sentinel);
    rc = b200sa_group_unbwt(gpus, work, n, sentinel);                        /* reverse_burrows_wheeler_transform, in place */
    if (rc) { printf("ERROR %d %s\n", rc, b200sa_last_error()); return 1; }
    printf("UNBWT %s\n", memcmp(work, text, (size_t)n) == 0 ? "roundtrip-ok" : "MISMATCH");
    if (n > 16) {                                                           /* corrupted input is rejected, the buffer stays as it was */
        memcpy(work, text, (size_t)n);
        b200sa_group_bwt(gpus, work, n, &sentinel);
        work[n / 2] ^= 1;
        memcpy(text, work, (size_t)n);
        rc = b200sa_group_unbwt(gpus, work, n, sentinel);
        printf("CORRUPT rc=%d unchanged=%d\n", rc, memcmp(work, text, (size_t)n) == 0);
    }
    b200sa_group_destroy(gpus);
    free(text); free(work); free(sa);
    return 0;
}
