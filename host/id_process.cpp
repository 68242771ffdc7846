#include "id_process.h"

#include <strings.h>

#include <cstdio>
#include <cstdlib>
#include <cstring>

namespace sqhost {

static const char *DELIMS = " !\"#$%&'()*+,-./:;<=>?@[\\]^_`{|}~";      // global `delimiters` @0x67a898 (.rodata+0x2be8)

// strSplit @0x40e0d0: maximal runs of non-delimiter characters
static void str_split(const std::string &s, std::vector<std::string> &out) {
    out.clear();
    size_t pos = s.find_first_not_of(DELIMS, 0), end = s.find_first_of(DELIMS, pos);
    while (end != std::string::npos || pos != std::string::npos) {
        out.push_back(s.substr(pos, end - pos));
        pos = s.find_first_not_of(DELIMS, end);
        end = s.find_first_of(DELIMS, pos);
    }
}
// strSplitWithDeli @0x430360: the same, also collecting the first delimiter after every token
static void str_split_deli(const std::string &s, std::vector<std::string> &out, std::string &deli) {
    out.clear(); deli.clear();
    size_t pos = s.find_first_not_of(DELIMS, 0), end = s.find_first_of(DELIMS, pos);
    while (pos != std::string::npos) {
        if (end != std::string::npos) deli.push_back(s[end]);
        out.push_back(s.substr(pos, end - pos));
        pos = s.find_first_not_of(DELIMS, end);
        end = s.find_first_of(DELIMS, pos);
    }
}
static bool all_digits(const std::string &s) {
    for (char c : s) if (c < '0' || c > '9') return false;
    return true;
}

// analysisPEType @0x430f50
static int pe_type(const std::string &a, const std::string &b) {
    if (a.size() == b.size() && memcmp(a.data(), b.data(), a.size()) == 0) return 1;
    // (the reference compares a's prefix against b without looking at b's length; an ID shorter than that prefix
    // cannot match, and is not read past its end here)
    if (!a.empty() && !b.empty() && a[a.size() - 1] == '1' && b[b.size() - 1] == '2' &&
        b.size() >= a.size() - 1 && a.compare(0, a.size() - 1, b, 0, a.size() - 1) == 0) return 2;
    size_t p = a.find("length");
    if (p != std::string::npos && (int)p != -1 && b.size() >= p && a.compare(0, p, b, 0, p) == 0) return 3;
    return 0;
}

bool analyse_id_bins(const MemBuf &mb, bool pe, uint8_t idbin[512]) {
    const uint32_t n = mb.nreads();
    if (n == 0) { idbin[0] = 1; return true; }          // (the reference never sees an empty sample block)
    const uint8_t *p = mb.ids.data();
    std::string first((const char *)p, mb.idlen[0]);
    std::vector<std::string> prev, cur;
    str_split(first, prev);
    int lenfield = -1;
    auto setbin = [&](size_t i, uint8_t v) { if (i < 512) idbin[i] = v; };
    if (!prev.empty()) {
        for (size_t i = 1;; i++) {                        // @0x4311ab: i runs 1 .. size
            if (lenfield >= 0) {
                if (lenfield == (int)i - 1) setbin(i + 1, 3);
            } else if (strcasecmp(prev[i - 1].c_str(), "length") == 0 || strcasecmp(prev[i - 1].c_str(), "len") == 0) {
                if (i < prev.size()) {
                    char *endp = nullptr;
                    long v = strtol(prev[i].c_str(), &endp, 10);
                    if (endp != prev[i].c_str() && (int32_t)v == mb.seqlen[0]) { setbin(i + 1, 0); lenfield = (int)i; }
                }
            }
            if (prev.size() <= i) break;
        }
    }
    bool seen[4] = {false, false, false, false};
    uint32_t i = 0;
    while (i < n) {
        std::string id1((const char *)p, mb.idlen[i]);
        p += mb.idlen[i];
        i++;
        if (pe) {
            if (i >= n) break;
            std::string id2((const char *)p, mb.idlen[i]);
            p += mb.idlen[i];
            i++;
            seen[pe_type(id1, id2)] = true;
        }
        str_split(id1, cur);
        if (cur.size() != prev.size()) { idbin[0] = 0; return false; }
        for (size_t j = 0; j < cur.size(); j++) {
            if ((j != 0 && (int)j - 1 == lenfield) || (int)j == lenfield) continue;
            if (cur[j].size() == prev[j].size() && memcmp(cur[j].data(), prev[j].data(), cur[j].size()) == 0) { setbin(j + 2, 0); continue; }
            if (!all_digits(prev[j]) || !all_digits(cur[j]) || prev[j].empty() || cur[j].empty()) { idbin[0] = 0; return false; }
            unsigned long a = strtoul(prev[j].c_str(), nullptr, 10), b = strtoul(cur[j].c_str(), nullptr, 10);
            if (b - a != 1) { idbin[0] = 0; return false; }
            setbin(j + 2, 1);
        }
        prev.swap(cur);
    }
    if (!pe) { idbin[0] = 1; return true; }
    if (seen[0]) { idbin[0] = 0; idbin[1] = 0; return false; }
    idbin[0] = 1;
    idbin[1] = seen[3] ? 3 : seen[2] ? 2 : seen[1] ? 1 : 0;
    return true;
}

struct IdBin { uint8_t type; unsigned long value; std::string str; };

// assembleNameS @0x4300a0
static void assemble(uint32_t r, bool pe, const std::vector<IdBin> &bins, const std::string &deli, const int32_t *seqlen, std::string &out) {
    out.clear();
    char tmp[32];
    for (size_t i = 0; i < bins.size(); i++) {
        const IdBin &b = bins[i];
        if (b.type == 1) { snprintf(tmp, sizeof tmp, "%lu", b.value + (pe ? r / 2 : r)); out += tmp; }
        else if (b.type == 3) { snprintf(tmp, sizeof tmp, "%d", seqlen[r]); out += tmp; }
        else if (b.type < 1) out += b.str;
        if (i + 1 < bins.size()) out.push_back(i < deli.size() ? deli[i] : '\0');
    }
}

bool decode_ids(const uint8_t *payload, size_t paylen, const uint8_t idbin[512], const int32_t *seqlen, uint32_t nreads,
                std::vector<uint8_t> &ids, std::vector<uint16_t> &idlen) {
    if (paylen < 2) return false;
    size_t len = payload[0] | (payload[1] << 8);
    if (2 + len > paylen || len >= 1024) return false;
    std::string first((const char *)payload + 2, strnlen((const char *)payload + 2, len));   // built from a zeroed C buffer
    std::vector<std::string> fields;
    std::string deli;
    str_split_deli(first, fields, deli);
    std::vector<IdBin> bins(fields.size());
    for (size_t i = 0; i < fields.size(); i++) {
        bins[i].type = i + 2 < 512 ? idbin[i + 2] : 0;
        if (bins[i].type == 1) bins[i].value = strtoul(fields[i].c_str(), nullptr, 10);
        else bins[i].str = fields[i].substr(0, 63);
    }
    const int petype = idbin[1];
    std::string name, second;
    auto put = [&](const std::string &s) {
        ids.insert(ids.end(), s.begin(), s.end());
        idlen.push_back((uint16_t)s.size());
    };
    if (petype == 0) {
        for (uint32_t r = 0; r < nreads; r++) { assemble(r, false, bins, deli, seqlen, name); put(name); }
        return true;
    }
    const char lastdeli = deli.empty() ? '\0' : deli[deli.size() - 1];
    for (uint32_t r = 0; 2 * r < nreads; r++) {
        assemble(2 * r, true, bins, deli, seqlen, name);
        put(name);
        if (2 * r + 1 >= nreads) break;
        if (petype == 2) { second = name; if (!second.empty()) second[second.size() - 1] = '2'; put(second); }
        else if (petype == 3) {
            size_t pos = name.rfind(lastdeli);
            second = name.substr(0, pos == std::string::npos ? 0 : pos + 1);
            char tmp[32];
            snprintf(tmp, sizeof tmp, "%d", seqlen[2 * r + 1]);
            second += tmp;
            put(second);
        } else if (petype == 1) put(name);
    }
    return true;
}

// ---- RFC 1321
static inline uint32_t rol(uint32_t x, int s) { return (x << s) | (x >> (32 - s)); }
static void md5_block(uint32_t h[4], const uint8_t *p) {
    static const uint32_t K[64] = {
        0xd76aa478, 0xe8c7b756, 0x242070db, 0xc1bdceee, 0xf57c0faf, 0x4787c62a, 0xa8304613, 0xfd469501, 0x698098d8, 0x8b44f7af, 0xffff5bb1,
        0x895cd7be, 0x6b901122, 0xfd987193, 0xa679438e, 0x49b40821, 0xf61e2562, 0xc040b340, 0x265e5a51, 0xe9b6c7aa, 0xd62f105d, 0x02441453,
        0xd8a1e681, 0xe7d3fbc8, 0x21e1cde6, 0xc33707d6, 0xf4d50d87, 0x455a14ed, 0xa9e3e905, 0xfcefa3f8, 0x676f02d9, 0x8d2a4c8a, 0xfffa3942,
        0x8771f681, 0x6d9d6122, 0xfde5380c, 0xa4beea44, 0x4bdecfa9, 0xf6bb4b60, 0xbebfbc70, 0x289b7ec6, 0xeaa127fa, 0xd4ef3085, 0x04881d05,
        0xd9d4d039, 0xe6db99e5, 0x1fa27cf8, 0xc4ac5665, 0xf4292244, 0x432aff97, 0xab9423a7, 0xfc93a039, 0x655b59c3, 0x8f0ccc92, 0xffeff47d,
        0x85845dd1, 0x6fa87e4f, 0xfe2ce6e0, 0xa3014314, 0x4e0811a1, 0xf7537e82, 0xbd3af235, 0x2ad7d2bb, 0xeb86d391};
    static const int S[64] = {7, 12, 17, 22, 7, 12, 17, 22, 7, 12, 17, 22, 7, 12, 17, 22, 5, 9, 14, 20, 5, 9, 14, 20, 5, 9, 14, 20, 5, 9, 14, 20,
                              4, 11, 16, 23, 4, 11, 16, 23, 4, 11, 16, 23, 4, 11, 16, 23, 6, 10, 15, 21, 6, 10, 15, 21, 6, 10, 15, 21, 6, 10, 15, 21};
    uint32_t w[16], a = h[0], b = h[1], c = h[2], d = h[3];
    for (int i = 0; i < 16; i++) w[i] = p[4 * i] | (p[4 * i + 1] << 8) | (p[4 * i + 2] << 16) | ((uint32_t)p[4 * i + 3] << 24);
    for (int i = 0; i < 64; i++) {
        uint32_t f;
        int g;
        if (i < 16) { f = (b & c) | (~b & d); g = i; }
        else if (i < 32) { f = (d & b) | (~d & c); g = (5 * i + 1) & 15; }
        else if (i < 48) { f = b ^ c ^ d; g = (3 * i + 5) & 15; }
        else { f = c ^ (b | ~d); g = (7 * i) & 15; }
        f += a + K[i] + w[g];
        a = d; d = c; c = b;
        b += rol(f, S[i]);
    }
    h[0] += a; h[1] += b; h[2] += c; h[3] += d;
}
void md5_digest(const uint8_t *data, size_t n, uint8_t out[16]) {
    uint32_t h[4] = {0x67452301, 0xefcdab89, 0x98badcfe, 0x10325476};
    size_t i = 0;
    for (; i + 64 <= n; i += 64) md5_block(h, data + i);
    uint8_t tail[128] = {0};
    size_t rem = n - i;
    memcpy(tail, data + i, rem);
    tail[rem] = 0x80;
    size_t tl = rem < 56 ? 64 : 128;
    uint64_t bits = (uint64_t)n * 8;
    for (int k = 0; k < 8; k++) tail[tl - 8 + k] = (uint8_t)(bits >> (8 * k));
    md5_block(h, tail);
    if (tl == 128) md5_block(h, tail + 64);
    for (int k = 0; k < 4; k++) { out[4 * k] = (uint8_t)h[k]; out[4 * k + 1] = (uint8_t)(h[k] >> 8); out[4 * k + 2] = (uint8_t)(h[k] >> 16); out[4 * k + 3] = (uint8_t)(h[k] >> 24); }
}

}  // namespace sqhost
