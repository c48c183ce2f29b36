// fmt.hpp — the few Rust formatting rules the reference's logs depend on (println! of f32/f64, {:?} of
// nested arrays, {:.5?} of the Genome structs) and the genome file format of main.rs:199-225.
#pragma once
#include <charconv>
#include <cstdint>
#include <cstdio>
#include <fstream>
#include <sstream>
#include <stdexcept>
#include <string>
#include <vector>

namespace evosops {

// Rust Display for floats: shortest representation that round-trips, never exponent form for these ranges
template <class F>
inline std::string fmt_display(F v) {
    char buf[64];
    auto r = std::to_chars(buf, buf + sizeof(buf), v, std::chars_format::fixed);
    std::string s(buf, r.ptr);
    // to_chars(fixed) without precision is shortest round-trip in fixed notation
    return s;
}
// Rust Debug for floats: Display plus ".0" when the value is integral
template <class F>
inline std::string fmt_debug(F v) {
    std::string s = fmt_display(v);
    if (s.find('.') == std::string::npos && s.find("inf") == std::string::npos && s.find("nan") == std::string::npos) s += ".0";
    return s;
}
inline std::string fmt_fixed(double v, int prec) {
    char buf[64];
    std::snprintf(buf, sizeof(buf), "%.*f", prec, v);
    return buf;
}

// {:?} of a nested u8 array with the given dims, e.g. [[[4, 0, 0, 1], [6, 1, 0, 0]], ...]
inline void fmt_nested(std::string& out, const uint8_t* g, const int* dims, int nd) {
    out += '[';
    if (nd == 1) {
        for (int k = 0; k < dims[0]; ++k) { if (k) out += ", "; out += std::to_string((unsigned)g[k]); }
    } else {
        int stride = 1;
        for (int k = 1; k < nd; ++k) stride *= dims[k];
        for (int k = 0; k < dims[0]; ++k) { if (k) out += ", "; fmt_nested(out, g + (size_t)k * stride, dims + 1, nd - 1); }
    }
    out += ']';
}
inline std::string fmt_genome_debug(const uint8_t* g, const int* dims, int nd) {
    std::string s;
    fmt_nested(s, g, dims, nd);
    return s;
}

// main.rs:199-210: strip '[', ']', ' ', pop() the last character unconditionally, split on ',',
// keep what parses as u8
inline std::vector<uint8_t> parse_genome_text(const std::string& contents) {
    std::string stripped;
    for (char c : contents) if (c != '[' && c != ']' && c != ' ') stripped += c;
    if (!stripped.empty()) stripped.pop_back();
    std::vector<uint8_t> out;
    std::stringstream ss(stripped);
    std::string tok;
    while (std::getline(ss, tok, ',')) {
        if (tok.empty()) continue;
        bool ok = true;
        unsigned v = 0;
        for (char c : tok) { if (c < '0' || c > '9') { ok = false; break; } v = v * 10 + (unsigned)(c - '0'); if (v > 255) { ok = false; break; } }
        if (ok) out.push_back((uint8_t)v);
    }
    return out;
}
inline std::vector<uint8_t> read_genome_file(const std::string& path) {
    std::ifstream f(path);
    if (!f) throw std::runtime_error("cannot read genome file " + path);
    std::stringstream ss;
    ss << f.rdbuf();
    return parse_genome_text(ss.str());
}

}  // namespace evosops
