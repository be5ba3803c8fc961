// On-disk formats of the circom toolchain around the witness path, host side:
//
//   .sym   text, one line per signal: `labelIdx,varIdx,componentIdx,name` (circom 0.5.x `-s`); varIdx is the witness
//          index the compiler gave the signal, -1 when constraint reduction (`-r`) eliminated it.  hz_load_sym turns it
//          into a layout map: for every witness index of the COMPILER's layout the index of the same signal in this
//          library's own layout -- applied as a gather when a witness leaves the device, so that the bytes of a `.wtns`
//          follow the compiler's numbering (and a reduced layout keeps only the signals the compiler kept).
//   .r1cs  binary (iden3 r1cs v1): header (field size, prime, nWires, nPubOut, nPubIn, nPrvIn, nLabels, mConstraints),
//          constraints as three sparse linear combinations A, B, C of (wire, coefficient) with A.w * B.w = C.w,
//          wire-to-label map.  hz_load_r1cs flattens the constraints into CSR arrays for the checker kernel
//          (k_r1cs_check) -- the independent sanity check `tester.checkConstraints` / `loadConstraints` give the
//          reference's tests (/root/reference/test/rollup-main.test.js:52-54).
#pragma once
#include <algorithm>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

static inline uint64_t hz_hash64(const char* s, size_t n) {           // FNV-1a, 64 bit, finished with a mix
    uint64_t h = 1469598103934665603ull;
    for (size_t i = 0; i < n; i++) { h ^= (uint8_t)s[i]; h *= 1099511628211ull; }
    h ^= h >> 29; h *= 0xbf58476d1ce4e5b9ull; h ^= h >> 32;
    return h;
}

struct HzLayoutMap {
    std::vector<uint32_t> src;      // compiler's witness index -> this library's witness index
    std::vector<uint32_t> inv;      // this library's index -> compiler's (0xffffffff: not in the compiler's witness)
    bool active() const { return !src.empty(); }
};

// names: (hash of the full name, own index) of every signal of the circuit, any order.  Returns 0, or a negative code:
// -1 unreadable file / malformed line, -2 a name of the file is not a signal of this circuit, -3 a witness index of the
// compiler's layout is left without a signal, -4 two names of this circuit collide in the 64-bit hash (cannot map safely)
static inline int hz_load_sym(const char* path, std::vector<std::pair<uint64_t, uint32_t> >& names, uint32_t own_nvars,
                              HzLayoutMap& M, std::string* why) {
    std::sort(names.begin(), names.end());
    for (size_t k = 1; k < names.size(); k++)
        if (names[k].first == names[k - 1].first && names[k].second != names[k - 1].second) { if (why) *why = "hash collision between two signal names"; return -4; }
    FILE* f = fopen(path, "r");
    if (!f) { if (why) *why = std::string("cannot open ") + path; return -1; }
    std::vector<uint32_t> src(1, 0u);              // witness entry 0 is the constant 1 in both layouts
    std::vector<uint8_t> have(1, 1);
    std::vector<uint32_t> inv(own_nvars, 0xffffffffu);
    if (own_nvars) inv[0] = 0;
    char line[1024];
    uint64_t ln = 0;
    int rc = 0;
    while (fgets(line, sizeof line, f)) {
        ln++;
        size_t n = strlen(line);
        while (n > 0 && (line[n - 1] == '\n' || line[n - 1] == '\r')) line[--n] = 0;
        if (n == 0) continue;
        // labelIdx,varIdx,componentIdx,name
        char* p1 = strchr(line, ','); if (!p1) { rc = -1; break; }
        char* p2 = strchr(p1 + 1, ','); if (!p2) { rc = -1; break; }
        char* p3 = strchr(p2 + 1, ','); if (!p3) { rc = -1; break; }
        const long long var = atoll(p1 + 1);
        if (var < 0) continue;                     // eliminated by the compiler's reduction
        const char* name = p3 + 1;
        const uint64_t h = hz_hash64(name, strlen(name));
        auto it = std::lower_bound(names.begin(), names.end(), std::make_pair(h, 0u));
        if (it == names.end() || it->first != h) { rc = -2; if (why) *why = std::string("not a signal of this circuit: ") + name; break; }
        if ((uint64_t)var >= src.size()) { src.resize((size_t)var + 1, 0xffffffffu); have.resize((size_t)var + 1, 0); }
        if (!have[(size_t)var]) { src[(size_t)var] = it->second; have[(size_t)var] = 1; }
        // signals the compiler joined may own separate entries here (and the other way round): each keeps its way back
        if (it->second < own_nvars && inv[it->second] == 0xffffffffu) inv[it->second] = (uint32_t)var;
    }
    fclose(f);
    if (rc == -1 && why) { char b[64]; snprintf(b, sizeof b, "malformed .sym line %llu", (unsigned long long)ln); *why = b; }
    if (rc) return rc;
    for (size_t v = 0; v < src.size(); v++)
        if (!have[v]) { if (why) { char b[96]; snprintf(b, sizeof b, "witness index %zu of the compiler's layout has no signal in the .sym", v); *why = b; } return -3; }
    M.src.swap(src);
    M.inv.swap(inv);
    return 0;
}

// ---- .r1cs -------------------------------------------------------------------------------------------------------
struct HzR1cs {
    uint32_t n_wires = 0, n_pub_out = 0, n_pub_in = 0, n_prv_in = 0, n_constraints = 0;
    uint64_t n_labels = 0;
    // CSR: the terms of constraint c are [off[3 c + k], off[3 c + k + 1]) for k = 0 (A), 1 (B), 2 (C)
    std::vector<uint64_t> off;
    std::vector<uint32_t> wire;
    std::vector<uint8_t> coef;      // 32 B little-endian canonical each
};
// 0, or -1 unreadable / malformed, -2 not the BN254 scalar field, -3 a wire id out of range
static inline int hz_load_r1cs(const char* path, const uint8_t prime_le[32], HzR1cs& R, std::string* why) {
    FILE* f = fopen(path, "rb");
    if (!f) { if (why) *why = std::string("cannot open ") + path; return -1; }
    auto fail = [&](int rc, const char* m) { if (why) *why = m; fclose(f); return rc; };
    uint8_t h[12];
    if (fread(h, 1, 12, f) != 12 || memcmp(h, "r1cs", 4) != 0) return fail(-1, "not an r1cs file");
    uint32_t version, nsec; memcpy(&version, h + 4, 4); memcpy(&nsec, h + 8, 4);
    if (version != 1) return fail(-1, "unsupported r1cs version");
    // sections may come in any order: find them first
    struct Sec { uint32_t type; uint64_t size; long pos; };
    std::vector<Sec> secs;
    for (uint32_t s = 0; s < nsec; s++) {
        uint8_t sh[12];
        if (fread(sh, 1, 12, f) != 12) return fail(-1, "truncated section header");
        Sec e; memcpy(&e.type, sh, 4); memcpy(&e.size, sh + 4, 8); e.pos = ftell(f);
        secs.push_back(e);
        if (fseek(f, (long)e.size, SEEK_CUR) != 0) return fail(-1, "truncated section");
    }
    const Sec* hdr = nullptr; const Sec* cons = nullptr;
    for (auto& e : secs) { if (e.type == 1) hdr = &e; if (e.type == 2) cons = &e; }
    if (!hdr || !cons) return fail(-1, "header or constraint section missing");
    fseek(f, hdr->pos, SEEK_SET);
    uint32_t fs;
    if (fread(&fs, 4, 1, f) != 1 || fs != 32) return fail(-2, "field size is not 32 bytes");
    uint8_t prime[32];
    if (fread(prime, 1, 32, f) != 32 || memcmp(prime, prime_le, 32) != 0) return fail(-2, "the prime is not the BN254 scalar field");
    uint32_t v[4];
    if (fread(v, 4, 4, f) != 4) return fail(-1, "truncated header");
    R.n_wires = v[0]; R.n_pub_out = v[1]; R.n_pub_in = v[2]; R.n_prv_in = v[3];
    if (fread(&R.n_labels, 8, 1, f) != 1 || fread(&R.n_constraints, 4, 1, f) != 1) return fail(-1, "truncated header");
    fseek(f, cons->pos, SEEK_SET);
    R.off.assign(1, 0);
    R.off.reserve(3ull * R.n_constraints + 1);
    for (uint32_t c = 0; c < R.n_constraints; c++)
        for (int k = 0; k < 3; k++) {
            uint32_t nt;
            if (fread(&nt, 4, 1, f) != 1) return fail(-1, "truncated constraint");
            const size_t at = R.wire.size();
            R.wire.resize(at + nt); R.coef.resize(32 * (at + nt));
            for (uint32_t t = 0; t < nt; t++) {
                if (fread(&R.wire[at + t], 4, 1, f) != 1 || fread(&R.coef[32 * (at + t)], 1, 32, f) != 32) return fail(-1, "truncated term");
                if (R.wire[at + t] >= R.n_wires) return fail(-3, "wire id out of range");
            }
            R.off.push_back(R.wire.size());
        }
    fclose(f);
    return 0;
}
