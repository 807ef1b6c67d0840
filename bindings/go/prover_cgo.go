// cgo binding of libzkcensus_b200 for go.vocdoni.io/dvote/crypto/zk/prover: replaces the body of prover.Prove
// (called at zk_census_test.go:89 of vocdoni/zk-franchise-proof-circuit).  Proof, ParseProof and (*Proof).Bytes stay
// as they are in that package.  NOT compiled in this repository's image (no Go toolchain); the same C entry points
// are exercised through ctypes in tests/test_gpu_prover.py.  See INTEGRATION.md section 1.
package prover

/*
#cgo LDFLAGS: -L${SRCDIR}/lib -lzkcensus_b200 -lcudart
#include <stdlib.h>
#include "zkcensus_b200.h"
*/
import "C"

import (
	"crypto/sha256"
	"fmt"
	"sync"
	"unsafe"
)

// A loaded circuit is found again in two steps.  The fast path is the identity of the caller's slices
// (data pointer + length of zkey and wasm): zk_census_test.go reads both files once and passes the same slices to
// every Prove, so the 55 MB key is NOT hashed per call (round 1 did, ~30 ms against a 6 ms proof).  A pair of slices
// seen for the first time is hashed once (sha256) and mapped to an already loaded circuit with the same contents.
type sliceID struct {
	zp, wp uintptr
	zn, wn int
}

var (
	mu       sync.Mutex
	ctx      *C.zkb_ctx
	byID     = map[sliceID]*C.zkb_circuit{}
	byDigest = map[[64]byte]*C.zkb_circuit{} // sha256(zkey) || sha256(wasm)
)

func circuitFor(zkey, wasm []byte) (*C.zkb_circuit, error) {
	if len(zkey) == 0 || len(wasm) == 0 {
		return nil, fmt.Errorf("empty zkey or wasm")
	}
	mu.Lock()
	defer mu.Unlock()
	if ctx == nil {
		if rc := C.zkb_ctx_create(0, &ctx); rc != 0 {
			return nil, fmt.Errorf("zkb_ctx_create: %s", C.GoString(C.zkb_last_error()))
		}
	}
	id := sliceID{uintptr(unsafe.Pointer(&zkey[0])), uintptr(unsafe.Pointer(&wasm[0])), len(zkey), len(wasm)}
	if c, ok := byID[id]; ok {
		return c, nil
	}
	var d [64]byte
	zs, ws := sha256.Sum256(zkey), sha256.Sum256(wasm)
	copy(d[:32], zs[:])
	copy(d[32:], ws[:])
	c, ok := byDigest[d]
	if !ok {
		rc := C.zkb_load_circuit(ctx, unsafe.Pointer(&zkey[0]), C.size_t(len(zkey)),
			unsafe.Pointer(&wasm[0]), C.size_t(len(wasm)), &c)
		if rc != 0 {
			return nil, fmt.Errorf("zkb_load_circuit (%d): %s", rc, C.GoString(C.zkb_last_error()))
		}
		byDigest[d] = c
	}
	if len(byID) > 64 { // slices come and go (GC); the digest map is the durable one
		byID = map[sliceID]*C.zkb_circuit{}
	}
	byID[id] = c
	return c, nil
}

// Prove keeps the signature used at zk_census_test.go:89.
func Prove(zkey, wasm, inputs []byte) (*Proof, error) {
	if len(inputs) == 0 {
		return nil, fmt.Errorf("empty inputs")
	}
	c, err := circuitFor(zkey, wasm)
	if err != nil {
		return nil, err
	}
	var info [8]C.uint32_t
	C.zkb_circuit_info(c, &info[0])
	pubCap := 96*int(info[1]) + 64 // public.json: up to 77 digits, quotes and a separator per public signal
	if pubCap < 2048 {
		pubCap = 2048
	}
	proofBuf := make([]byte, 1024)
	pubBuf := make([]byte, pubCap)
	errBuf := make([]byte, 256)
	pn, qn := C.size_t(len(proofBuf)), C.size_t(len(pubBuf))
	rc := C.zkb_fullprove(c, (*C.char)(unsafe.Pointer(&inputs[0])), C.size_t(len(inputs)),
		(*C.char)(unsafe.Pointer(&proofBuf[0])), &pn, (*C.char)(unsafe.Pointer(&pubBuf[0])), &qn,
		(*C.char)(unsafe.Pointer(&errBuf[0])), C.size_t(len(errBuf)))
	if rc != 0 { // 4 = circuit assert failed (the wasm's exceptionHandler(4))
		return nil, fmt.Errorf("prove (%d): %s", rc, C.GoString((*C.char)(unsafe.Pointer(&errBuf[0]))))
	}
	return ParseProof(proofBuf[:pn], pubBuf[:qn]) // unchanged: zk_census_test.go:118
}

// Verify runs the Groth16 pairing check on the GPU ((*Proof).Verify at zk_census_test.go:122).
func Verify(vkey, pubSignals, proof []byte) error {
	if len(vkey) == 0 || len(pubSignals) == 0 || len(proof) == 0 {
		return fmt.Errorf("empty verification key, public signals or proof")
	}
	rc := C.zkb_verify((*C.char)(unsafe.Pointer(&vkey[0])), C.size_t(len(vkey)),
		(*C.char)(unsafe.Pointer(&pubSignals[0])), C.size_t(len(pubSignals)),
		(*C.char)(unsafe.Pointer(&proof[0])), C.size_t(len(proof)))
	if rc != 0 {
		return fmt.Errorf("verify (%d): %s", rc, C.GoString(C.zkb_last_error()))
	}
	return nil
}
