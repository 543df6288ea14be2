// kxpu_cgo.go -- the cgo shim of INTEGRATION.md as a file: drop it into pkg/device_plugin of the
// reference next to device_plugin.go (build tag keeps the stock CPU path selectable).
// NOT compiled in this repository: the build image has no Go toolchain.
//go:build kxpu

package device_plugin

/*
#cgo CFLAGS: -I${SRCDIR}/../../include
#cgo LDFLAGS: -L${SRCDIR}/../../lib -lkxpu
#include <stdlib.h>
#include "kxpu.h"
*/
import "C"

import (
	"fmt"
	"os"
	"runtime"
	"unsafe"
)

type kxpu struct {
	ctx   *C.kxpu_ctx
	table *C.kxpu_table // parsed pci.ids, built once
}

func kxCheck(ctx *C.kxpu_ctx, what string, rc C.int32_t) error {
	if rc == C.KXPU_OK {
		return nil
	}
	return fmt.Errorf("%s: %s (%s)", what, C.GoString(C.kxpu_strerror(rc)), C.GoString(C.kxpu_last_error(ctx)))
}

// The library has no CPU path: without a GPU on the NVIDIA driver kxpu_ctx_create returns
// KXPU_E_NOGPU.  A node whose GPUs are ALL bound to vfio-pci (the reference's normal deployment)
// therefore cannot use it; InitiateDevicePlugin calls newKxpu once and, on error, logs it and keeps
// the stock functions (createIommuDeviceMap, getDeviceName, generateCDISpec are untouched in the
// tree and selected by `if kx == nil`).  That is the host's choice between two implementations,
// not a fallback inside the library; after a successful create every kxpu_* error is fatal.
func newKxpu(ordinal int) (*kxpu, error) {
	k := &kxpu{}
	if rc := C.kxpu_ctx_create(C.int32_t(ordinal), &k.ctx); rc != C.KXPU_OK {
		return nil, fmt.Errorf("kxpu_ctx_create: %s", C.GoString(C.kxpu_strerror(rc)))
	}
	return k, nil
}

// S2: parse pciIdsFilePath once (replaces the per-id rescan of getDeviceName/locateVendor).
func (k *kxpu) loadPciIds(path string) error {
	text, err := os.ReadFile(path)
	if err != nil {
		return err
	}
	var p *C.uint8_t
	if len(text) > 0 {
		p = (*C.uint8_t)(unsafe.Pointer(&text[0])) // borrowed for the call only
	}
	return kxCheck(k.ctx, "kxpu_pciids_load", C.kxpu_pciids_load(k.ctx, p, C.size_t(len(text)), &k.table))
}

// S2: batched getDeviceName for NVIDIA device ids ("" = not found, like the reference).
func (k *kxpu) deviceNames(ids []string) ([]string, error) {
	keys := make([]C.uint32_t, len(ids))
	for i, id := range ids {
		var d uint32
		if _, err := fmt.Sscanf(id, "%04x", &d); err != nil || len(id) != 4 {
			keys[i] = 0xffffffff // never present under 10de -> ""
			continue
		}
		keys[i] = C.uint32_t(0x10de<<16 | d)
	}
	rows := make([]C.int32_t, len(ids))
	offs := make([]C.uint32_t, len(ids)+1)
	if len(ids) == 0 {
		return nil, nil
	}
	if err := kxCheck(k.ctx, "kxpu_lookup", C.kxpu_lookup(k.ctx, k.table, &keys[0], C.size_t(len(ids)), &rows[0])); err != nil {
		return nil, err
	}
	var need C.size_t
	C.kxpu_names(k.ctx, k.table, &rows[0], C.size_t(len(ids)), nil, 0, &offs[0], &need) // sizing call
	buf := make([]byte, need+1)
	if err := kxCheck(k.ctx, "kxpu_names", C.kxpu_names(k.ctx, k.table, &rows[0], C.size_t(len(ids)),
		(*C.uint8_t)(unsafe.Pointer(&buf[0])), need, &offs[0], &need)); err != nil {
		return nil, err
	}
	out := make([]string, len(ids))
	for i := range ids {
		out[i] = string(buf[offs[i]:offs[i+1]])
	}
	return out, nil
}

// S1/S4: classify the raw sysfs records gathered by the walk.
// kxpu_classify_out is a struct of seven output pointers.  `out` itself lives in Go memory, so cgo
// (cgocheck=1, the default) inspects it and refuses Go pointers to UNPINNED Go memory inside it
// ("cgo argument has Go pointer to unpinned Go pointer").  runtime.Pinner (Go 1.21+) pins the seven
// backing arrays for the duration of the call; the library keeps none of them (include/kxpu.h).
func (k *kxpu) classify(recs []C.kxpu_devrec) (accept, gids, goff, gmem []uint32, dids []uint64, doff, dgrp []uint32, err error) {
	n := len(recs)
	accept, gids, goff, gmem = make([]uint32, n), make([]uint32, n), make([]uint32, n+1), make([]uint32, n)
	dids, doff, dgrp = make([]uint64, n), make([]uint32, n+1), make([]uint32, n)
	if n == 0 {
		return
	}
	var pin runtime.Pinner
	defer pin.Unpin()
	pin.Pin(&accept[0])
	pin.Pin(&gids[0])
	pin.Pin(&goff[0])
	pin.Pin(&gmem[0])
	pin.Pin(&dids[0])
	pin.Pin(&doff[0])
	pin.Pin(&dgrp[0])
	var out C.kxpu_classify_out
	out.accept_index = (*C.uint32_t)(unsafe.Pointer(&accept[0]))
	out.group_ids = (*C.uint32_t)(unsafe.Pointer(&gids[0]))
	out.group_off = (*C.uint32_t)(unsafe.Pointer(&goff[0]))
	out.group_members = (*C.uint32_t)(unsafe.Pointer(&gmem[0]))
	out.dev_ids = (*C.uint64_t)(unsafe.Pointer(&dids[0]))
	out.dev_off = (*C.uint32_t)(unsafe.Pointer(&doff[0]))
	out.dev_groups = (*C.uint32_t)(unsafe.Pointer(&dgrp[0]))
	err = kxCheck(k.ctx, "kxpu_classify", C.kxpu_classify(k.ctx, &recs[0], C.size_t(n), &out))
	gids, goff, gmem = gids[:out.n_groups], goff[:out.n_groups+1], gmem[:out.n_accepted]
	dids, doff, dgrp = dids[:out.n_devids], doff[:out.n_devids+1], dgrp[:out.n_groups]
	return
}

// S2, start-up form: pci.ids text + every device id of deviceMap in ONE call and one host round trip
// (kxpu_pciids_join: for the 1.4 MB file one cooperative kernel parses, folds, names and joins).
func (k *kxpu) loadAndJoin(path string, keys []uint32) ([]int32, error) {
	text, err := os.ReadFile(path)
	if err != nil {
		return nil, err
	}
	rows := make([]int32, len(keys))
	var tp *C.uint8_t
	var kp *C.uint32_t
	var rp *C.int32_t
	if len(text) > 0 {
		tp = (*C.uint8_t)(unsafe.Pointer(&text[0]))
	}
	if len(keys) > 0 {
		kp = (*C.uint32_t)(unsafe.Pointer(&keys[0]))
		rp = (*C.int32_t)(unsafe.Pointer(&rows[0]))
	}
	err = kxCheck(k.ctx, "kxpu_pciids_join", C.kxpu_pciids_join(k.ctx, tp, C.size_t(len(text)), kp, C.size_t(len(keys)), rp, &k.table))
	return rows, err
}

// S3: CDI document bytes (format 0 = YAML as the reference's live path, 1 = JSON).
func (k *kxpu) cdiEmit(format int, devs []C.kxpu_cdidev) ([]byte, error) {
	var p *C.kxpu_cdidev
	if len(devs) > 0 {
		p = &devs[0]
	}
	var n C.size_t
	C.kxpu_cdi_emit(k.ctx, C.int32_t(format), p, C.size_t(len(devs)), nil, 0, &n) // sizing call
	buf := make([]byte, n+1)
	err := kxCheck(k.ctx, "kxpu_cdi_emit", C.kxpu_cdi_emit(k.ctx, C.int32_t(format), p, C.size_t(len(devs)),
		(*C.uint8_t)(unsafe.Pointer(&buf[0])), n, &n))
	return buf[:n], err
}

// S5: CDI device names of an Allocate response.
func (k *kxpu) allocNames(idx []uint64) ([]string, error) {
	if len(idx) == 0 {
		return nil, nil
	}
	offs := make([]C.uint32_t, len(idx)+1)
	buf := make([]byte, 36*len(idx))
	var need C.size_t
	err := kxCheck(k.ctx, "kxpu_alloc_names", C.kxpu_alloc_names(k.ctx, (*C.uint64_t)(unsafe.Pointer(&idx[0])), C.size_t(len(idx)),
		(*C.uint8_t)(unsafe.Pointer(&buf[0])), C.size_t(len(buf)), &offs[0], &need))
	out := make([]string, len(idx))
	for i := range idx {
		out[i] = string(buf[offs[i]:offs[i+1]])
	}
	return out, err
}
