/* Plain-C consumer of include/fadtk_b200_io.h (and a compile check of include/fadtk_b200.h): proves the drop-in
 * boundary is a C ABI - no C++ types, no Python in the loop.  Built and run by tests/test_fileio.py.
 * usage: io_abi_check <dir>   -> writes two WAVs and two .npy files into <dir>, reads them back, prints "ok". */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "../../include/fadtk_b200.h"
#include "../../include/fadtk_b200_io.h"

int main(int argc, char** argv) {
    if (argc != 2) return 2;
    char p0[512], p1[512], q0[512], q1[512];
    snprintf(p0, sizeof p0, "%s/a.wav", argv[1]);
    snprintf(p1, sizeof p1, "%s/b.wav", argv[1]);
    snprintf(q0, sizeof q0, "%s/a.npy", argv[1]);
    snprintf(q1, sizeof q1, "%s/b.npy", argv[1]);
    const char* wavs[2] = {p0, p1};
    const char* npys[2] = {q0, q1};
    int16_t pcm[300];
    for (int i = 0; i < 300; ++i) pcm[i] = (int16_t)(i * 7 - 1000);
    long long off[2] = {0, 100}, frames[2] = {100, 200};
    int status[2];
    if (fad_io_wav_write(wavs, 2, 2, pcm, off, frames, 16000, status) != 0) return 3;

    int sr[2], ch[2];
    long long fr[2];
    if (fad_io_wav_probe(wavs, 2, 0, sr, ch, fr, status) != 0) return 4;
    if (sr[0] != 16000 || ch[1] != 1 || fr[0] != 100 || fr[1] != 200) return 5;
    int16_t back[300];
    if (fad_io_wav_read(wavs, 2, 1, back, off, fr, ch, status) != 0) return 6;
    if (memcmp(back, pcm, sizeof pcm) != 0) return 7;

    uint16_t emb[6 * 4];                                       /* fp16 bit patterns, [6, 4] */
    for (int i = 0; i < 24; ++i) emb[i] = (uint16_t)(0x3c00 + i);
    long long roff[2] = {0, 2}, rows[2] = {2, 4};
    if (fad_io_npy_write_f16(npys, 2, 2, emb, roff, rows, 4, status) != 0) return 8;
    long long r[2];
    int c[2], nd[2], dt[2];
    if (fad_io_npy_probe(npys, 2, 2, r, c, nd, dt, status) != 0) return 9;
    if (r[0] != 2 || r[1] != 4 || c[0] != 4 || nd[1] != 2 || dt[0] != 2) return 10;
    uint16_t eback[24];
    if (fad_io_npy_read_f16(npys, 2, 2, eback, roff, r, 4, status) != 0) return 11;
    if (memcmp(eback, emb, sizeof emb) != 0) return 12;

    const char* missing[1] = {"/nonexistent/dir/x.wav"};
    if (fad_io_wav_probe(missing, 1, 1, sr, ch, fr, status) != 1 || status[0] != FAD_IO_EOPEN) return 13;
    printf("ok v%d\n", fad_io_version());
    return 0;
}
