/* A C host on the C ABI of include/hzwitness.h with nothing else underneath (no Python, no torch): the shape of the
 * reference's production call `./circuit input.json witness.wtns` (/root/reference/tools/helpers/actions.js:139-143),
 * with the input already packed (hzw_ninputs() x 32 B little-endian, declaration order; `python -m circuits_b200.cli`
 * is the JSON front end).
 *
 *   hzw_cli <rollup-main|withdraw|rollup-tx> <nTx> <nLevels> <maxL1Tx> <maxFeeTx> <inputs.bin> <witness.wtns> [sanity] [n_gpus]
 *
 * Prints `nvars=<n> w1=<hex of witness entry 1, big-endian>`; a violated constraint prints the runtime's message on
 * stderr and exits 1 like the reference's binary.
 *   gcc -O2 -I include bindings/c/hzw_cli.c circuits_b200/libhzwitness.so -Wl,-rpath,$PWD/circuits_b200 -o hzw_cli */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "hzwitness.h"

static void print_dec(FILE* f, const uint8_t le[32]) {
    uint32_t t[8];
    char buf[80];
    int n = 0, nz = 1;
    for (int i = 0; i < 8; i++) t[i] = (uint32_t)le[4 * i] | ((uint32_t)le[4 * i + 1] << 8) | ((uint32_t)le[4 * i + 2] << 16) | ((uint32_t)le[4 * i + 3] << 24);
    while (nz) {
        uint64_t rem = 0;
        nz = 0;
        for (int i = 7; i >= 0; i--) { uint64_t cur = (rem << 32) | t[i]; t[i] = (uint32_t)(cur / 10); rem = cur % 10; if (t[i]) nz = 1; }
        buf[n++] = (char)('0' + rem);
    }
    while (n > 0) fputc(buf[--n], f);
}

int main(int argc, char** argv) {
    if (argc < 8) {
        fprintf(stderr, "usage: %s <rollup-main|withdraw|rollup-tx> nTx nLevels maxL1Tx maxFeeTx inputs.bin witness.wtns [sanity] [n_gpus]\n", argv[0]);
        return 2;
    }
    hzw_params p;
    memset(&p, 0, sizeof p);
    p.circuit_kind = !strcmp(argv[1], "withdraw") ? HZW_CIRCUIT_WITHDRAW : (!strcmp(argv[1], "rollup-tx") ? HZW_CIRCUIT_ROLLUP_TX : HZW_CIRCUIT_ROLLUP_MAIN);
    p.nTx = atoi(argv[2]); p.nLevels = atoi(argv[3]); p.maxL1Tx = atoi(argv[4]); p.maxFeeTx = atoi(argv[5]);
    p.device = 0; p.n_gpus = argc > 9 ? atoi(argv[9]) : 1;
    const int sanity = argc > 8 ? atoi(argv[8]) : 0;
    hzw_ctx* ctx = NULL;
    int st = hzw_create(&p, NULL, &ctx);
    if (st != HZW_OK) { fprintf(stderr, "hzw_create failed with code %d%s\n", st, st == HZW_E_DEVICE ? " (no CUDA device: there is no CPU path)" : ""); return 3; }
    const uint64_t nin = hzw_ninputs(ctx), nvars = hzw_nvars(ctx);
    uint8_t* in = (uint8_t*)malloc(32 * nin);
    uint8_t* w = (uint8_t*)malloc(32 * nvars);
    FILE* f = fopen(argv[6], "rb");
    if (!in || !w || !f || fread(in, 32, nin, f) != nin) { fprintf(stderr, "cannot read %llu inputs from %s\n", (unsigned long long)nin, argv[6]); return 2; }
    fclose(f);
    hzw_error err;
    st = hzw_calculate(ctx, in, sanity, w, &err);
    if (st == HZW_E_CONSTRAINT) {
        fprintf(stderr, "Constraint doesn't match ");
        print_dec(stderr, err.lhs); fprintf(stderr, " != "); print_dec(stderr, err.rhs);
        fprintf(stderr, " %s\n", err.where);
        return 1;
    }
    if (st != HZW_OK) { fprintf(stderr, "hzw_calculate failed with code %d\n", st); return 3; }
    if (hzw_write_wtns(ctx, w, argv[7]) != HZW_OK) { fprintf(stderr, "cannot write %s\n", argv[7]); return 2; }
    printf("nvars=%llu w1=", (unsigned long long)nvars);
    for (int i = 31; i >= 0; i--) printf("%02x", w[32 + i]);
    printf("\n");
    free(in); free(w);
    hzw_destroy(ctx);
    return 0;
}
