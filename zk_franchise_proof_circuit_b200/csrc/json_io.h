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
bool dec_to_fq(const char *s, size_t len, uint32_t out[8]);   // mod q
// object of numeric values (flattened, mod q when base_field else mod r); text values are skipped
bool parse_json_numbers(const char *s, size_t len, bool base_field, std::map<std::string, std::vector<uint32_t>> &out,
                        std::string &err);
// true when any value parsed on this thread since json_reset_reduced() was outside [0, modulus) - verifiers must
// reject such encodings instead of silently reducing them
bool json_value_was_reduced();
void json_reset_reduced();
// array of decimal strings / numbers, flattened, mod r
bool parse_json_array(const char *s, size_t len, std::vector<uint32_t> &out, std::string &err);

// proof256 = A.x A.y | B.x.c0 B.x.c1 B.y.c0 B.y.c1 | C.x C.y, canonical LE.  snarkjs_style adds protocol/curve.
std::string proof_to_json(const uint8_t *proof256, bool snarkjs_style);
std::string publics_to_json(const uint8_t *pub, uint32_t n);

}  // namespace zkb
