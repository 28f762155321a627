package tokenmonster

/*
#cgo LDFLAGS: -L${SRCDIR}/../../tokenmonster_b200 -ltmgpu
#cgo CFLAGS: -I${SRCDIR}/../../include
#include <stdlib.h>
#include "tmgpu.h"
*/
import "C"

import (
	"errors"
	"unsafe"
)

// GPUVocab is a vocabulary resident on one B200.  Create it from the same file Load reads.
type GPUVocab struct{ h *C.tmgpu_vocab }

func lastErr(rc C.int) error {
	return errors.New("tmgpu " + string(rune('0'+rc)) + ": " + C.GoString(C.tmgpu_last_error()))
}

// LoadGPU is Load (go/tokenmonster.go:2656) for the GPU path.
func LoadGPU(filename string, device int) (*GPUVocab, error) {
	cs := C.CString(filename)
	defer C.free(unsafe.Pointer(cs))
	var h *C.tmgpu_vocab
	if rc := C.tmgpu_vocab_load_file(cs, C.int(device), &h); rc != 0 {
		return nil, lastErr(rc)
	}
	return &GPUVocab{h}, nil
}

func (v *GPUVocab) Close() { C.tmgpu_vocab_destroy(v.h); v.h = nil }

// flatten: cgo may not pass [][]byte; one blob + offsets, nothing is retained after return.
func flatten(batch [][]byte) ([]byte, []C.uint64_t) {
	total := 0
	for _, b := range batch {
		total += len(b)
	}
	blob := make([]byte, 0, total+1)
	off := make([]C.uint64_t, len(batch)+1)
	for i, b := range batch {
		off[i] = C.uint64_t(len(blob))
		blob = append(blob, b...)
	}
	off[len(batch)] = C.uint64_t(len(blob))
	return blob, off
}

// TokenizeBatch is Vocab.Tokenize (go:959) for every string of the batch.
// errs[i] != nil  <=>  the reference would have returned an error for string i (go:963-966).
func (v *GPUVocab) TokenizeBatch(batch [][]byte) (tokens [][]uint32, missing []int, errs []error, err error) {
	blob, off := flatten(batch)
	var data *C.uint8_t
	if len(blob) > 0 {
		data = (*C.uint8_t)(unsafe.Pointer(&blob[0]))
	}
	var res *C.tmgpu_result
	if rc := C.tmgpu_tokenize_batch(v.h, data, &off[0], C.uint32_t(len(batch)), 4, &res); rc != 0 {
		return nil, nil, nil, lastErr(rc)
	}
	defer C.tmgpu_result_free(res)
	n := len(batch)
	ids := unsafe.Slice((*uint32)(unsafe.Pointer(res.tokens)), int(res.n_tokens))
	toff := unsafe.Slice((*uint64)(unsafe.Pointer(res.tok_off)), n+1)
	miss := unsafe.Slice((*uint32)(unsafe.Pointer(res.missing)), n)
	stat := unsafe.Slice((*uint8)(unsafe.Pointer(res.status)), n)
	tokens, missing, errs = make([][]uint32, n), make([]int, n), make([]error, n)
	for i := 0; i < n; i++ {
		tokens[i] = append([]uint32(nil), ids[toff[i]:toff[i+1]]...) // copy out of pinned memory
		missing[i] = int(miss[i])
		if stat[i] != 0 {
			errs[i] = errors.New("normalization failed") // server status 13
		}
	}
	return
}

// TokenizeToSerializedBatch is Vocab.TokenizeToSerialized (go:986); encodingLength 0, 2, 3 or 4.
func (v *GPUVocab) TokenizeToSerializedBatch(batch [][]byte, encodingLength uint8) (out [][]byte, used uint8, missing []int, err error) {
	blob, off := flatten(batch)
	var data *C.uint8_t
	if len(blob) > 0 {
		data = (*C.uint8_t)(unsafe.Pointer(&blob[0]))
	}
	var res *C.tmgpu_result
	if rc := C.tmgpu_tokenize_batch(v.h, data, &off[0], C.uint32_t(len(batch)), C.uint8_t(encodingLength), &res); rc != 0 {
		return nil, 0, nil, lastErr(rc)
	}
	defer C.tmgpu_result_free(res)
	n, w := len(batch), uint64(res.enc_len)
	raw := unsafe.Slice((*byte)(unsafe.Pointer(res.tokens)), int(uint64(res.n_tokens)*w))
	toff := unsafe.Slice((*uint64)(unsafe.Pointer(res.tok_off)), n+1)
	miss := unsafe.Slice((*uint32)(unsafe.Pointer(res.missing)), n)
	out, missing = make([][]byte, n), make([]int, n)
	for i := 0; i < n; i++ {
		out[i] = append([]byte(nil), raw[toff[i]*w:toff[i+1]*w]...)
		missing[i] = int(miss[i])
	}
	return out, uint8(w), missing, nil
}

// CountBatch is Vocab.Count (go:971).
func (v *GPUVocab) CountBatch(batch [][]byte) (counts []int, missing []int, err error) {
	blob, off := flatten(batch)
	n := len(batch)
	c := make([]C.uint64_t, n+1)
	m := make([]C.uint32_t, n+1)
	var data *C.uint8_t
	if len(blob) > 0 {
		data = (*C.uint8_t)(unsafe.Pointer(&blob[0]))
	}
	if rc := C.tmgpu_count_batch(v.h, data, &off[0], C.uint32_t(n), &c[0], &m[0], nil); rc != 0 {
		return nil, nil, lastErr(rc)
	}
	counts, missing = make([]int, n), make([]int, n)
	for i := 0; i < n; i++ {
		counts[i], missing[i] = int(c[i]), int(m[i])
	}
	return
}

// DecodeSerializedBatch is Vocab.DecodeSerialized (go:803) for every element; encodingLength 2, 3 or 4.
func (v *GPUVocab) DecodeSerializedBatch(batch [][]byte, encodingLength uint8) ([][]byte, error) {
	blob, off := flatten(batch)
	w := C.uint64_t(encodingLength)
	for i := range off {
		off[i] /= w // offsets are in ids
	}
	var data *C.uint8_t
	if len(blob) > 0 {
		data = (*C.uint8_t)(unsafe.Pointer(&blob[0]))
	}
	var res *C.tmgpu_bytes
	if rc := C.tmgpu_decode_batch(v.h, data, &off[0], C.uint32_t(len(batch)), C.uint8_t(encodingLength), &res); rc != 0 {
		return nil, lastErr(rc)
	}
	defer C.tmgpu_bytes_free(res)
	n := len(batch)
	raw := unsafe.Slice((*byte)(unsafe.Pointer(res.bytes)), int(res.n_bytes))
	boff := unsafe.Slice((*uint64)(unsafe.Pointer(res.off)), n+1)
	out := make([][]byte, n)
	for i := 0; i < n; i++ {
		out[i] = append([]byte(nil), raw[boff[i]:boff[i+1]]...)
	}
	return out, nil
}
