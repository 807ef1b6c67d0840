/* Driver for the transpiled reference witness calculator (TEST INFRASTRUCTURE, not product).
 *
 * Speaks the circom_runtime protocol that go-rapidsnark/witness and snarkjs use against
 * circuit.wasm: init(sanity) -> for every input value
 * writeSharedRWMemory x8 + setInputSignal(hashMSB, hashLSB, index) -> getWitness(i) +
 * readSharedRWMemory x8.  The wasm `runtime.exceptionHandler(code)` import is turned into a
 * longjmp so that a failed circuit assert (code 4) is returned to the caller instead of aborting.
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <setjmp.h>
#include "wasm_rt.h"

#define MAX_PAGES 4096u /* 256 MiB arena */
uint8_t *wasm_mem = 0;
uint32_t wasm_pages = 0;
extern const uint32_t wasm_min_pages;
static uint8_t *image = 0;
static size_t image_len = 0;
static jmp_buf trap_env;
static int trap_code = 0;

uint32_t w2c_getFieldNumLen32(void);
uint32_t w2c_getWitnessSize(void);
uint32_t w2c_getInputSize(void);
uint32_t w2c_readSharedRWMemory(uint32_t);
void w2c_writeSharedRWMemory(uint32_t, uint32_t);
void w2c_init(uint32_t);
void w2c_setInputSignal(uint32_t, uint32_t, uint32_t);
void w2c_getWitness(uint32_t);
void w2c_getRawPrime(void);
uint32_t w2c_getInputSignalSize(uint32_t, uint32_t);

uint32_t wasm_grow(uint32_t n) {
  uint32_t old = wasm_pages;
  if (wasm_pages + n > MAX_PAGES) return (uint32_t)-1;
  wasm_pages += n;
  return old;
}
void wasm_trap(void) { trap_code = 100; longjmp(trap_env, 1); }
void imp_exceptionHandler(uint32_t code) { trap_code = (int)code; longjmp(trap_env, 1); }
void imp_printErrorMessage(void) {}
void imp_writeBufferMessage(void) {}
void imp_showSharedRWMemory(void) {}

int wc_load(const char *mem_path) {
  FILE *f = fopen(mem_path, "rb");
  if (!f) return -1;
  fseek(f, 0, SEEK_END);
  image_len = (size_t)ftell(f);
  fseek(f, 0, SEEK_SET);
  image = (uint8_t *)malloc(image_len);
  if (fread(image, 1, image_len, f) != image_len) { fclose(f); return -2; }
  fclose(f);
  if (!wasm_mem) wasm_mem = (uint8_t *)calloc((size_t)MAX_PAGES, 65536);
  return wasm_mem ? 0 : -3;
}

uint32_t wc_witness_size(void) { return w2c_getWitnessSize(); }
uint32_t wc_input_size(void) { return w2c_getInputSize(); }

/* returns 0 on success, else the circom exception code (4 = constraint assert failed). */
int wc_witness(int n, const uint64_t *name_hash, const uint32_t *idx, const uint32_t *vals8,
               uint8_t *out, int sanity) {
  size_t used = (size_t)wasm_pages * 65536;
  if (used > image_len) memset(wasm_mem + image_len, 0, used - image_len);
  memcpy(wasm_mem, image, image_len);
  wasm_pages = wasm_min_pages;
  trap_code = 0;
  if (setjmp(trap_env)) return trap_code;
  w2c_init((uint32_t)sanity);
  for (int i = 0; i < n; i++) {
    for (uint32_t j = 0; j < 8; j++) w2c_writeSharedRWMemory(j, vals8[i * 8 + j]);
    w2c_setInputSignal((uint32_t)(name_hash[i] >> 32), (uint32_t)name_hash[i], idx[i]);
  }
  uint32_t nw = w2c_getWitnessSize();
  for (uint32_t i = 0; i < nw; i++) {
    w2c_getWitness(i);
    for (uint32_t j = 0; j < 8; j++) {
      uint32_t v = w2c_readSharedRWMemory(j);
      memcpy(out + (size_t)i * 32 + j * 4, &v, 4);
    }
  }
  return 0;
}
