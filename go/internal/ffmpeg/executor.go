//go:build cgo

// Package ffmpeg — B200 executor for the VCP consumer.
//
// Drop-in for the two functions the consumer calls today
// (cmd/consumer.go:370 runFFmpegWithTimeout, cmd/consumer.go:396 verifyOutputFile): same
// arguments, same error strings, but the work happens in libvcpenc (hand-written sm_100a CUDA
// kernels) instead of an `ffmpeg` child process.  B200 has no NVENC block, so the reference's
// default h264-nvenc* presets cannot run there at all.
//
// NOTE: the build image for this repository has no Go toolchain; this file is the binding a
// maintainer adds to the reference tree (see INTEGRATION.md).  It is deliberately tiny: every
// decision lives behind the C ABI in include/vcpenc.h.
//
// Build: CGO_ENABLED=1 (the reference Makefile uses CGO_ENABLED=0, Makefile:25), with
//   CGO_CFLAGS=-I<repo>/include  CGO_LDFLAGS="-L<repo>/video_codec_pipeline_b200/lib -lvcpenc"
package ffmpeg

/*
#cgo LDFLAGS: -lvcpenc
#include <stdlib.h>
#include "vcpenc.h"
*/
import "C"

import (
	"context"
	"errors"
	"fmt"
	"os"
	"os/exec"
	"strings"
	"sync/atomic"
	"time"
	"unsafe"
)

// ErrHandOff is returned for tasks that are not a B200 video encode: presets without a video
// encode (`-c copy`, `-vn`: VCPENC_E_NOTENCODE), inputs whose audio the loaded FFmpeg libraries
// cannot carry into the output (VCPENC_E_AUDIO) and video tools the library does not implement
// (VCPENC_E_UNSUPPORTED).  The caller runs those through RunStockFFmpeg -- see TranscodeOrStock.
// That is dispatch on the kind of task, not a CPU fallback of the encoder.
var ErrHandOff = errors.New("task is handed to the stock ffmpeg")

// ErrNotEncode is kept for callers of the first version of this binding.
var ErrNotEncode = ErrHandOff

// TranscodeOrStock is what cmd/consumer.go:262 calls: the B200 encoder, and the reference's own
// ffmpeg child for the tasks it hands off.
func TranscodeOrStock(ctx context.Context, input, output, ffmpegArgs string, timeout time.Duration) error {
	err := Transcode(ctx, input, output, ffmpegArgs, timeout)
	if errors.Is(err, ErrHandOff) {
		return RunStockFFmpeg(ctx, input, output, ffmpegArgs, timeout)
	}
	return err
}

// Transcode mirrors runFFmpegWithTimeout(ctx, input, output, ffmpegArgs, timeout).
func Transcode(ctx context.Context, input, output, ffmpegArgs string, timeout time.Duration) error {
	timeoutCtx, cancel := context.WithTimeout(ctx, timeout)
	defer cancel()

	fields := strings.Fields(ffmpegArgs) // whitespace split only, as cmd/consumer.go:378
	argv := make([]*C.char, len(fields)+1)
	for i, f := range fields {
		argv[i] = C.CString(f)
		defer C.free(unsafe.Pointer(argv[i]))
	}
	cIn, cOut := C.CString(input), C.CString(output)
	defer C.free(unsafe.Pointer(cIn))
	defer C.free(unsafe.Pointer(cOut))

	// exec.CommandContext SIGKILLs the child on cancel; a cgo call cannot be killed, so the
	// library polls this flag once per GOP while reading and between the stages of a chunk.
	var flag int32
	done := make(chan struct{})
	go func() {
		select {
		case <-timeoutCtx.Done():
			atomic.StoreInt32(&flag, 1)
		case <-done:
		}
	}()
	errbuf := make([]byte, 1024)
	rc := C.vcpenc_transcode(cIn, cOut, C.int(len(fields)), (**C.char)(unsafe.Pointer(&argv[0])),
		C.int(timeout/time.Millisecond), (*C.int)(unsafe.Pointer(&flag)),
		(*C.char)(unsafe.Pointer(&errbuf[0])), C.size_t(len(errbuf)))
	close(done)

	if timeoutCtx.Err() == context.DeadlineExceeded || rc == C.VCPENC_E_TIMEOUT {
		return fmt.Errorf("编码超时 (>%s)", timeout) // cmd/consumer.go:388
	}
	if ctx.Err() != nil || rc == C.VCPENC_E_CANCELLED {
		return fmt.Errorf("任务被取消") // cmd/consumer.go:391
	}
	switch rc {
	case C.VCPENC_OK:
		return nil
	case C.VCPENC_E_NOTENCODE, C.VCPENC_E_AUDIO, C.VCPENC_E_UNSUPPORTED:
		return fmt.Errorf("%w: %s", ErrHandOff, C.GoString((*C.char)(unsafe.Pointer(&errbuf[0]))))
	default:
		return fmt.Errorf("vcpenc error %d: %s", int(rc), C.GoString((*C.char)(unsafe.Pointer(&errbuf[0]))))
	}
}

// Verify mirrors verifyOutputFile(path) (cmd/consumer.go:396-419).
func Verify(path string) error {
	cPath := C.CString(path)
	defer C.free(unsafe.Pointer(cPath))
	errbuf := make([]byte, 512)
	if rc := C.vcpenc_verify(cPath, (*C.char)(unsafe.Pointer(&errbuf[0])), C.size_t(len(errbuf))); rc != 0 {
		return errors.New(C.GoString((*C.char)(unsafe.Pointer(&errbuf[0]))))
	}
	return nil
}

// RunStockFFmpeg is the reference's original body (cmd/consumer.go:370-394), kept for the tasks Transcode hands off.
func RunStockFFmpeg(ctx context.Context, input, output, ffmpegArgs string, timeout time.Duration) error {
	timeoutCtx, cancel := context.WithTimeout(ctx, timeout)
	defer cancel()
	args := []string{"-hide_banner", "-loglevel", "warning", "-y", "-i", input}
	if ffmpegArgs != "" {
		args = append(args, strings.Fields(ffmpegArgs)...)
	}
	args = append(args, output)
	cmd := exec.CommandContext(timeoutCtx, "ffmpeg", args...)
	cmd.Stdout, cmd.Stderr = os.Stdout, os.Stderr
	err := cmd.Run()
	if timeoutCtx.Err() == context.DeadlineExceeded {
		return fmt.Errorf("编码超时 (>%s)", timeout)
	}
	if ctx.Err() != nil {
		return fmt.Errorf("任务被取消")
	}
	return err
}

// DeviceCount honours CUDA_VISIBLE_DEVICES (install.sh:279-297 starts one consumer per GPU).
func DeviceCount() int { return int(C.vcpenc_device_count()) }
