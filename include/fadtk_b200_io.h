/* fadtk_b200_io.h - batched, multi-threaded file I/O either side of the GPU hot path (host only, no CUDA).
 *
 * The reference moves every clip through three per-file Python round trips:
 *   torchaudio.load / torchaudio.save of <dir>/convert/<sr>/<stem>.wav      fadtk/fad.py:139-160
 *   np.save(<dir>/embeddings/<model>/<stem>.npy, fp16 [n_frames, d])         fadtk/fad.py:188-201
 *   np.load of the same files                                                fadtk/fad.py:203-209, utils.py:13-16
 * With the forward passes at 10^4..10^5 x real time those round trips - GIL-bound at < 1000 files/s - are the
 * end-to-end limit of the directory flow.  These entry points do the same byte-compatible I/O for a whole batch
 * of files on native threads: plain pointers and sizes, caller-owned buffers (typically one pinned host buffer
 * that is then copied to the GPU in one piece), one status code per file, no exceptions.
 *
 * Conventions: `paths` = n NUL-terminated UTF-8 paths; `threads` <= 0 picks min(n, hardware threads, 32);
 * every function returns the number of files whose status is non-zero, or -1 for bad arguments.
 * Parent directories must exist (the caller creates each directory once, not once per file).
 */
#ifndef FADTK_B200_IO_H
#define FADTK_B200_IO_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define FAD_IO_OK 0
#define FAD_IO_EOPEN 1        /* cannot open / create */
#define FAD_IO_EFORMAT 2      /* not a RIFF/WAVE or .npy file, or a malformed header */
#define FAD_IO_EUNSUPPORTED 3 /* valid file, but not 16-bit integer PCM / not a C-ordered 1-D or 2-D array of the asked dtype */
#define FAD_IO_ESHORT 4       /* fewer bytes than the header (or the caller) promised; short write */

int fad_io_version(void);

/* WAV headers: sample rate, channel count and frames (samples per channel) of PCM16 files; what
 * ModelLoader.load_wav (model_loader.py:63-70) and FrechetAudioDistance.load_audio (fad.py:139-160) read. */
int fad_io_wav_probe(const char* const* paths, int n, int threads, int* sample_rate, int* channels,
                     long long* frames, int* status);

/* Samples of n PCM16 files into dst: file i occupies dst[offsets[i] .. offsets[i] + frames[i] * channels[i])
 * (interleaved as stored), with frames/channels as returned by fad_io_wav_probe. */
int fad_io_wav_read(const char* const* paths, int n, int threads, int16_t* dst, const long long* offsets,
                    const long long* frames, const int* channels, int* status);

/* Mono PCM_S16 RIFF files (the convert cache, fad.py:160): file i = src[offsets[i] .. offsets[i] + frames[i]). */
int fad_io_wav_write(const char* const* paths, int n, int threads, const int16_t* src, const long long* offsets,
                     const long long* frames, int sample_rate, int* status);

/* NumPy .npy (format 1.0, as np.save writes it) fp16 [rows[i], d] arrays - the embedding cache (fad.py:200):
 * file i = src rows row_offsets[i] .. row_offsets[i] + rows[i] of a row-major fp16 [*, d] buffer. */
int fad_io_npy_write_f16(const char* const* paths, int n, int threads, const void* src, const long long* row_offsets,
                         const long long* rows, int d, int* status);

/* Shape and dtype of .npy files: rows, cols (1-D arrays report cols = 1 and ndim = 1), itemsize-coded dtype
 * (2 = '<f2', 4 = '<f4', 8 = '<f8'; anything else -> FAD_IO_EUNSUPPORTED). */
int fad_io_npy_probe(const char* const* paths, int n, int threads, long long* rows, int* cols, int* ndim,
                     int* dtype_code, int* status);

/* Payload of fp16 .npy files with `d` columns into one row-major buffer (ragged concatenation):
 * file i -> dst rows row_offsets[i] .. row_offsets[i] + rows[i]. */
int fad_io_npy_read_f16(const char* const* paths, int n, int threads, void* dst, const long long* row_offsets,
                        const long long* rows, int d, int* status);

#ifdef __cplusplus
}
#endif
#endif
