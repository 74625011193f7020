"""BASELINE.json config 5 at a length the round's GPU budget allows: ONE long-form stream (goforward.raw repeated,
2.8 minutes by default) through audio -> features -> senone scores -> fwdtree + fwdflat on the device at the beams
1e-48 / 1e-64 / 1e-80, next to the compiled reference decoding the same samples with the same beam on the host
(ps_decode_raw path of oracle/ref_driver.c, one core).  Prints one JSON object."""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import refdrv                                   # noqa: E402  (the CPU arm: test / bench infrastructure)
from pocketsphinx_b200.decoder import Decoder               # noqa: E402

REF = os.path.join(ROOT, "oracle", "_ref")


def main():
    reps = int(sys.argv[1]) if len(sys.argv) > 1 else 60
    big = len(sys.argv) > 2 and sys.argv[2] == "lvcsr"
    hd = os.path.join(REF, "model", "en-us")
    dic = os.path.join(REF, "model", "cmudict-en-us.dict") if big else os.path.join(REF, "data", "turtle.dic")
    lm = os.path.join(REF, "model", "en-us.lm.bin") if big else os.path.join(REF, "data", "turtle.lm.bin")
    go = np.fromfile(os.path.join(REF, "data", "goforward.raw"), np.int16)
    pcm = np.tile(go, reps)
    out = {"stream_s": len(pcm) / 16000.0, "lm": os.path.basename(lm), "dict": os.path.basename(dic), "beams": {}}
    for beam in ("1e-48", "1e-64", "1e-80"):
        dec = Decoder(hd, dic, lm, max_utts=1, max_frames=len(pcm) // 160 + 16, beam=beam)
        dec.decode_raw_batch([go])                          # warm-up: allocations, first launches
        t0 = time.perf_counter()
        o = dec.decode_raw_batch([pcm])[0]
        t1 = time.perf_counter()
        dec.close()
        r = {"frames": int(o["n_frames"]), "words": len(o["hyp"].split()), "device_decode_s": t1 - t0,
             "device_frames_per_s": o["n_frames"] / (t1 - t0), "device_xrt": (t1 - t0) / (len(pcm) / 16000.0)}
        if os.environ.get("NO_REF") != "1":
            t2 = time.perf_counter()
            want = refdrv.decode(hd, lm, dic, pcm, bestpath="no", compallsen="yes", beam=beam)
            t3 = time.perf_counter()
            r.update(reference_decode_s=t3 - t2, reference_frames_per_s=o["n_frames"] / (t3 - t2),
                     same_words=o["hyp"] == want["hyp"], score=int(o["score"]), reference_score=int(want["score"]))
        out["beams"][beam] = r
        print(beam, r, file=sys.stderr, flush=True)
    print(json.dumps(out))


if __name__ == "__main__":
    main()
