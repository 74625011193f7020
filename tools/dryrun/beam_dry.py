# Dry run (CPU) of the config-5 measurement's logic: one long stream (goforward.raw repeated) through the Decoder's own
# code at several -beam values, device stages replaced as in decoder_dry.py; every hypothesis / score must equal a plain
# reference decode with the same beam.  Usage: python tools/dryrun/beam_dry.py [repeats]
import os, sys, time
import numpy as np
src = open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "decoder_dry.py")).read()
exec(src[:src.index("go = np.fromfile")])            # the stubs and the patched decoder module

reps = int(sys.argv[1]) if len(sys.argv) > 1 else 8
if len(sys.argv) > 2 and sys.argv[2] == "lvcsr":      # the shipped 72 k-word dictionary and trigram LM
    DIC, LM = os.path.join(REF, "model", "cmudict-en-us.dict"), os.path.join(REF, "model", "en-us.lm.bin")
go = np.fromfile(os.path.join(REF, "data", "goforward.raw"), np.int16)
pcm = np.tile(go, reps)
for beam in ("1e-48", "1e-64", "1e-80"):
    dec = decoder.Decoder(HD, DIC, LM, beam=beam)
    t0 = time.time()
    out = dec.decode_raw_batch([pcm])[0]
    t1 = time.time()
    dec.close()
    want = refdrv.decode(HD, LM, DIC, pcm, bestpath="no", compallsen="yes", beam=beam)
    t2 = time.time()
    print("beam %s: %d frames, %d words, same hyp %s, same score %s (emulation %.1f s, reference %.1f s)" % (
        beam, out["n_frames"], len(out["hyp"].split()), out["hyp"] == want["hyp"], out["score"] == want["score"], t1 - t0, t2 - t1))
    assert out["hyp"] == want["hyp"] and out["score"] == want["score"]
print("beam dry run ok")
