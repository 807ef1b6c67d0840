/* Runtime shim for the C produced by oracle/wasm2c.py (TEST INFRASTRUCTURE, not product). */
#ifndef WASM_RT_H
#define WASM_RT_H
#include <stdint.h>
extern uint8_t *wasm_mem;
extern uint32_t wasm_pages;
extern void *wasm_table[];
uint32_t wasm_grow(uint32_t n);
void wasm_trap(void);
#endif
