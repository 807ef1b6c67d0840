// Minimal JSON <-> field-element helpers for the reference's file formats - see json_io.cc
#pragma once
#include <cstdint>
#include <map>
#include <string>
#include <vector>

namespace zkb {

// Parses an inputs.json object: {"name": "123" | 123 | ["1", ["2", ...]], ...}.  Values are flattened in
// order and reduced mod r to canonical 8 x u32 little-endian limbs.
bool parse_inputs_json(const char *s, size_t len, std::map<std::string, std::vector<uint32_t>> &out, std::string &err);

std::string u256_to_dec(const uint32_t v[8]);
bool dec_to_fr(const char *s, size_t len, uint32_t out[8]);   // mod r

// proof256 = A.x A.y | B.x.c0 B.x.c1 B.y.c0 B.y.c1 | C.x C.y, canonical LE.  snarkjs_style adds protocol/curve.
std::string proof_to_json(const uint8_t *proof256, bool snarkjs_style);
std::string publics_to_json(const uint8_t *pub, uint32_t n);

}  // namespace zkb
