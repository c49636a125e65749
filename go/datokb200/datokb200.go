// Package datokb200 is the Go side of the drop-in boundary: a type that implements
// datok.Tokenizer (fomafile.go:29-33) on top of libdatok_b200.so, the B200-native
// implementation of MatrixTokenizer.TransduceTokenWriter (matrix.go:348-698).
//
// NOTE: this file cannot be compiled in the build image (no Go toolchain); it is the
// binding a Datok maintainer would add.  Everything it needs from the native side is
// declared in include/datok_b200.h and exercised through the same C ABI by the Python
// host mirror (datok_b200/tokenizer.py) and the -m gpu parity tests.
package datokb200

/*
#cgo CFLAGS: -I${SRCDIR}/../../include
#cgo LDFLAGS: -L${SRCDIR}/../../datok_b200 -ldatok_b200
#include <stdlib.h>
#include "datok_b200.h"

// cgo cannot call Go closures from C directly; the replay goes through exported trampolines (//export
// below).  cgo declares exported functions without `const` in _cgo_export.h, so the prototypes here
// match that and the pointers are cast where the callback table is filled.
extern void goDatokToken(void *user, uint8_t *buf, size_t bufBytes, size_t offBytes, int32_t offRunes);
extern void goDatokSentenceEnd(void *user);
extern void goDatokTextEnd(void *user);

static int datok_replay_go(const datok_result *r, const uint8_t *in, size_t n, void *user) {
  datok_callbacks cb;
  cb.user = user;
  cb.token = (void (*)(void *, const uint8_t *, size_t, size_t, int32_t))goDatokToken;
  cb.sentence_end = goDatokSentenceEnd;
  cb.text_end = goDatokTextEnd;
  return datok_replay(r, in, n, &cb);
}
*/
import "C"

import (
	"errors"
	"io"
	"log"
	"os"
	"runtime/cgo"
	"unsafe"

	datok "github.com/KorAP/datok"
)

// Tokenizer implements datok.Tokenizer for .matok (and .datok) models on one B200.
type Tokenizer struct {
	model *C.datok_model
}

// compile-time check: same interface as the reference (fomafile.go:29-33)
var _ datok.Tokenizer = (*Tokenizer)(nil)

// BlockBytes is how much of the reader one push hands to the GPU.  The reference keeps a window of
// at most 1024 runes (matrix.go:365); here memory is bounded by one block plus the text that is still
// open at its end.
var BlockBytes = 64 << 20

// LoadTokenizerFile mirrors datok.LoadTokenizerFile (fomafile.go:452-484): nil on any error, the
// reason is logged.
func LoadTokenizerFile(file string, device int) *Tokenizer {
	cs := C.CString(file)
	defer C.free(unsafe.Pointer(cs))
	var rc C.int
	m := C.datok_load(cs, C.int(device), &rc)
	if m == nil {
		log.Println(C.GoString(C.datok_last_error()))
		return nil
	}
	return &Tokenizer{model: m}
}

// Automaton stands for what datok.LoadFomaFile returns (fomafile.go:36-52): a foma file waiting for
// ToMatrix.  The parsed intermediate representation lives inside the library.
type Automaton struct {
	file   string
	device int
}

// LoadFomaFile mirrors datok.LoadFomaFile (fomafile.go:56-72).
func LoadFomaFile(file string, device int) *Automaton {
	f, err := os.Open(file)
	if err != nil {
		log.Print(err)
		return nil
	}
	f.Close()
	return &Automaton{file: file, device: device}
}

// ToMatrix mirrors Automaton.ToMatrix (matrix.go:30-99): ParseFoma and the matrix build run in the
// library (datok_load_foma), the model is resident on the device when this returns.
func (a *Automaton) ToMatrix() *Tokenizer {
	cs := C.CString(a.file)
	defer C.free(unsafe.Pointer(cs))
	var rc C.int
	m := C.datok_load_foma(cs, C.int(a.device), &rc)
	if m == nil {
		log.Println(C.GoString(C.datok_last_error()))
		return nil
	}
	return &Tokenizer{model: m}
}

// Save mirrors MatrixTokenizer.Save (matrix.go:107-123): the gzipped WriteTo image.
func (t *Tokenizer) Save(file string) (n int64, err error) {
	cs := C.CString(file)
	defer C.free(unsafe.Pointer(cs))
	if rc := C.datok_save(t.model, cs); rc != C.DATOK_OK {
		return 0, errors.New(C.GoString(C.datok_last_error()))
	}
	return int64(C.datok_write_image(t.model, nil, 0)), nil
}

// WriteTo mirrors MatrixTokenizer.WriteTo (matrix.go:126-210).
func (t *Tokenizer) WriteTo(w io.Writer) (n int64, err error) {
	size := C.datok_write_image(t.model, nil, 0)
	if size == 0 {
		return 0, errors.New(C.GoString(C.datok_last_error()))
	}
	buf := make([]byte, int(size))
	C.datok_write_image(t.model, (*C.uint8_t)(unsafe.Pointer(&buf[0])), size)
	k, err := w.Write(buf)
	return int64(k), err
}

// Convert is `datok convert -i fomaFile -o matokFile` (cmd/datok.go:63) without touching a device.
func Convert(fomaFile, matokFile string) error {
	a, b := C.CString(fomaFile), C.CString(matokFile)
	defer C.free(unsafe.Pointer(a))
	defer C.free(unsafe.Pointer(b))
	if rc := C.datok_compile_foma(a, b); rc != C.DATOK_OK {
		return errors.New(C.GoString(C.datok_last_error()))
	}
	return nil
}

// Close releases the GPU resident model.
func (t *Tokenizer) Close() { C.datok_free(t.model); t.model = nil }

// Type is "MATOK" (matrix.go:102-104) or "DATOK" (datok.go:252-254), by the file's magic.
func (t *Tokenizer) Type() string { return C.GoString(C.datok_model_type(t.model)) }

func check(rc C.int) bool {
	if rc == C.DATOK_OK {
		return true
	}
	if rc <= C.DATOK_ERR_DEGENERATE {
		panic(errors.New(C.GoString(C.datok_strerror(rc)))) // the reference panics here too
	}
	log.Println(C.GoString(C.datok_last_error()))
	return false
}

// stream pushes the reader through datok_stream_* block by block (cmd/datok.go:108-132: any io.Reader,
// STDIN included) and hands every batch result, with the input bytes it covers, to deliver.
func (t *Tokenizer) stream(r io.Reader, flags C.uint32_t, deliver func(res *C.datok_result, in []byte)) bool {
	st := C.datok_stream_open(t.model, flags)
	if st == nil {
		return false
	}
	defer C.datok_stream_close(st)
	block := make([]byte, BlockBytes)
	var pending []byte // bytes pushed but not covered by a result yet (only kept for the replay)
	for {
		n, err := io.ReadFull(r, block)
		if n > 0 {
			var res *C.datok_result
			done0 := uint64(C.datok_stream_bytes_done(st))
			if !check(C.datok_stream_push(st, (*C.uint8_t)(unsafe.Pointer(&block[0])), C.size_t(n), &res)) {
				return false
			}
			pending = append(pending, block[:n]...)
			if res != nil {
				k := uint64(C.datok_stream_bytes_done(st)) - done0
				deliver(res, pending[:k])
				C.datok_result_free(res)
				pending = append(pending[:0], pending[k:]...)
			}
		}
		if err == io.EOF || err == io.ErrUnexpectedEOF {
			break
		}
		if err != nil {
			log.Fatalln(err) // matrix.go:401
			return false
		}
	}
	var res *C.datok_result
	if !check(C.datok_stream_finish(st, &res)) {
		return false
	}
	if res != nil {
		deliver(res, pending)
		C.datok_result_free(res)
	}
	return true
}

// Transduce mirrors matrix.go:340-342.  The stock SIMPLE writer's text is formatted on the device
// (DATOK_FORMAT) and written as it comes back.
func (t *Tokenizer) Transduce(r io.Reader, w io.Writer) bool {
	ok := t.stream(r, C.uint32_t(C.DATOK_TOKENS|C.DATOK_SENTENCES|C.DATOK_FORMAT), func(res *C.datok_result, _ []byte) {
		v := C.datok_result_view(res)
		if v.text_len > 0 {
			w.Write(unsafe.Slice((*byte)(unsafe.Pointer(v.text)), int(v.text_len)))
		}
	})
	return ok
}

// TransduceFlags is Transduce for any flag set of NewTokenWriter (token_writer.go:17-25): the text the
// stock writer would produce, formatted on the device.
func (t *Tokenizer) TransduceFlags(r io.Reader, w io.Writer, flags datok.Bits) bool {
	return t.stream(r, C.uint32_t(flags)|C.uint32_t(C.DATOK_FORMAT), func(res *C.datok_result, _ []byte) {
		v := C.datok_result_view(res)
		if v.text_len > 0 {
			w.Write(unsafe.Slice((*byte)(unsafe.Pointer(v.text)), int(v.text_len)))
		}
	})
}

// TransduceTokenWriter mirrors matrix.go:348-698: the input is transduced on the GPU
// and the event stream (Token / SentenceEnd / TextEnd, in the reference's order and
// with the reference's arguments) is replayed into the caller's TokenWriter, so custom
// writers (token_writer.go:27-33) keep working.  Inputs on which the reference panics
// make this function panic with the same cause.
func (t *Tokenizer) TransduceTokenWriter(r io.Reader, w *datok.TokenWriter) bool {
	defer w.Flush() // matrix.go:374
	h := cgo.NewHandle(w)
	defer h.Delete()
	// DATOK_COMPACT8: the replay only needs the delta-coded spans (4 bytes per token over PCIe)
	return t.stream(r, C.uint32_t(C.DATOK_TOKENS|C.DATOK_SENTENCES|C.DATOK_COMPACT8), func(res *C.datok_result, in []byte) {
		var p *C.uint8_t
		if len(in) > 0 {
			p = (*C.uint8_t)(unsafe.Pointer(&in[0]))
		}
		C.datok_replay_go(res, p, C.size_t(len(in)), unsafe.Pointer(&h))
	})
}

//export goDatokToken
func goDatokToken(user unsafe.Pointer, buf *C.uint8_t, bufBytes C.size_t, offBytes C.size_t, offRunes C.int32_t) {
	w := (*cgo.Handle)(user).Value().(*datok.TokenWriter)
	// []rune(string(bytes)) decodes exactly like bufio.Reader.ReadRune: one U+FFFD per bad byte
	runes := []rune(string(C.GoBytes(unsafe.Pointer(buf), C.int(bufBytes))))
	w.Token(int(offRunes), runes)
}

//export goDatokSentenceEnd
func goDatokSentenceEnd(user unsafe.Pointer) {
	(*cgo.Handle)(user).Value().(*datok.TokenWriter).SentenceEnd(0)
}

//export goDatokTextEnd
func goDatokTextEnd(user unsafe.Pointer) {
	(*cgo.Handle)(user).Value().(*datok.TokenWriter).TextEnd(0)
}

// Offsets is the array-level result for callers that do not need the closure replay:
// the TokenWriter's pos / sent lists and the token byte spans, straight from the GPU.
type Offsets struct {
	TokBytes []uint32 // 2 per token: surface = in[TokBytes[2k]:TokBytes[2k+1]]
	TokPos   []int32  // 2 per token: text-relative rune offsets (TokenWriter.pos)
	SentPos  []int32  // TokenWriter.sent entries
	TextTok  []uint32 // per TextEnd: tokens emitted so far
}

// TransduceOffsets runs the path and copies the offset arrays out of the pinned result.
func (t *Tokenizer) TransduceOffsets(in []byte, flags datok.Bits) (*Offsets, error) {
	var p *C.uint8_t
	if len(in) > 0 {
		p = (*C.uint8_t)(unsafe.Pointer(&in[0]))
	}
	var res *C.datok_result
	if rc := C.datok_transduce(t.model, p, C.size_t(len(in)), C.uint32_t(flags), nil, &res); rc != C.DATOK_OK {
		return nil, errors.New(C.GoString(C.datok_strerror(rc)))
	}
	defer C.datok_result_free(res)
	v := C.datok_result_view(res)
	nt, ns, nx := int(v.n_tokens), int(v.n_sent_pos), int(v.n_texts)
	o := &Offsets{}
	if v.tok_bytes != nil {
		o.TokBytes = append(o.TokBytes, unsafe.Slice((*uint32)(unsafe.Pointer(v.tok_bytes)), 2*nt)...)
	}
	if v.tok_pos != nil {
		o.TokPos = append(o.TokPos, unsafe.Slice((*int32)(unsafe.Pointer(v.tok_pos)), 2*nt)...)
	}
	if v.sent_pos != nil {
		o.SentPos = append(o.SentPos, unsafe.Slice((*int32)(unsafe.Pointer(v.sent_pos)), ns)...)
	}
	o.TextTok = append(o.TextTok, unsafe.Slice((*uint32)(unsafe.Pointer(v.text_tok_end)), nx)...)
	return o, nil
}
