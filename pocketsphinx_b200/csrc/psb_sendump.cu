// psb_sendump.cu -- the reference's senone-dump (".sen") wire format at the GMM <-> search
// boundary (acmod_write_senfh_header / acmod_write_scores / acmod_read_scores,
// acmod.c:335-346, 880-1017), host code only.  A batch scored on the GPU can be
// written in this format and replayed by the UNMODIFIED reference search through
// ps_decode_senscr (pocketsphinx.c:1200) or `pocketsphinx_batch -senin yes`.
#include "psb_internal.cuh"

#include <stdio.h>
#include <string.h>

#define PSB_BYTE_ORDER_MAGIC 0x11223344u      /* bio.h:91 */

// Text header "s3\nversion 0.1\nmdef_file <path>\nn_sen <N>\nlogbase <%f>\nendhdr\n" + magic,
// then per frame int16 n_active (= n_sen: all senones) followed by int16 scores[n_sen].
extern "C" int psb_sendump_write(const char *path, const char *mdef_file, int32_t n_sen, double logbase,
                                 const int16_t *senscr, int64_t n_frames)
{
    PSB_REQUIRE(path && senscr && n_sen > 0 && n_sen < 32768 && n_frames >= 0, "psb_sendump_write: bad argument");
    FILE *fp = fopen(path, "wb");
    PSB_REQUIRE(fp, "psb_sendump_write: cannot open %s", path);
    fprintf(fp, "s3\nversion 0.1\nmdef_file %s\nn_sen %d\nlogbase %f\nendhdr\n", mdef_file ? mdef_file : "(null)", n_sen,
            logbase);
    const uint32_t magic = PSB_BYTE_ORDER_MAGIC;
    bool ok = fwrite(&magic, 4, 1, fp) == 1;
    const int16_t n16 = (int16_t)n_sen;
    for (int64_t t = 0; ok && t < n_frames; ++t)
        ok = fwrite(&n16, 2, 1, fp) == 1 && fwrite(senscr + t * n_sen, 2, n_sen, fp) == (size_t)n_sen;
    ok = fclose(fp) == 0 && ok;
    PSB_REQUIRE(ok, "psb_sendump_write: write to %s failed", path);
    return PSB_OK;
}

// Reads a dump written by the reference (`-senlogdir`) or by psb_sendump_write.  Frames with
// fewer than n_sen active senones are expanded like acmod_read_scores_internal does
// (unlisted senones get SENSCR_DUMMY = 0x7fff, acmod.h:60).  Returns the number of frames read
// (<= max_frames), or a negative psb_status_t.
extern "C" int64_t psb_sendump_read(const char *path, int32_t *n_sen_out, int16_t *senscr, int64_t max_frames)
{
    PSB_REQUIRE(path, "psb_sendump_read: null path");
    FILE *fp = fopen(path, "rb");
    PSB_REQUIRE(fp, "psb_sendump_read: cannot open %s", path);
    char line[1024];
    int n_sen = -1;
    bool hdr = false;
    while (fgets(line, sizeof(line), fp)) {
        if (!strncmp(line, "endhdr", 6)) { hdr = true; break; }
        if (!strncmp(line, "n_sen ", 6)) n_sen = atoi(line + 6);
    }
    uint32_t magic = 0;
    if (!hdr || n_sen <= 0 || fread(&magic, 4, 1, fp) != 1 || magic != PSB_BYTE_ORDER_MAGIC) {
        fclose(fp);
        psb_set_error("psb_sendump_read: %s is not a native-endian senone dump", path);
        return PSB_ERR_ARG;
    }
    if (n_sen_out) *n_sen_out = n_sen;
    int64_t t = 0;
    std::vector<uint8_t> deltas(n_sen);
    while (senscr && t < max_frames) {
        int16_t n_active;
        if (fread(&n_active, 2, 1, fp) != 1) break;
        int16_t *row = senscr + t * n_sen;
        if (n_active == n_sen) {
            if (fread(row, 2, n_sen, fp) != (size_t)n_sen) break;
        }
        else {
            if (n_active < 0 || n_active > n_sen || fread(deltas.data(), 1, n_active, fp) != (size_t)n_active) break;
            for (int i = 0; i < n_sen; ++i) row[i] = 0x7fff;
            int n = 0;
            bool good = true;
            for (int i = 0; i < n_active && good; ++i) {
                n += deltas[i];
                good = n < n_sen && fread(row + n, 2, 1, fp) == 1;
            }
            if (!good) break;
        }
        ++t;
    }
    fclose(fp);
    return t;
}
