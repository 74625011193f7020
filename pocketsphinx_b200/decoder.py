"""Batch n-gram decoding from audio to words on the device, configured like the reference's ps_decoder_t
(-hmm, -dict, -lm and the search settings by their reference names): front end -> senone scores -> phone
loop -> first pass -> second pass on the GPU, hypothesis and segments read from the returned tables
(ps_decode_raw + ps_get_hyp / ps_seg_iter with -bestpath no, pocketsphinx.c:1073-1345).  Everything the
reference loads from files is read here by the package itself (s3io, lmio, dict2pid, lextree).

STATUS: GPU-verified end to end (tests/test_gpu_zz_decoder.py, first hardware run in round 2).  Out of the
hot-path scope: kept small, not extended.
"""
import math
import os

import numpy as np

from . import api, lextree
from .fe_tables import make_fe_desc
from .model import PackedModel

# config_macro.h (the reference's defaults for every front-end / feature key this class looks at)
FE_REFERENCE_DEFAULTS = dict(feat="1s_c_d_dd", cmn="live", agc="none", varnorm="no", lda="", svspec="", dither="no",
                             round_filters="yes", ncep="13", frate="100", nfft="0", cmninit="40,3,-1", samprate="16000",
                             wlen="0.025625", nfilt="40", lowerf="133.33334", upperf="6855.4976", alpha="0.97",
                             transform="legacy", lifter="0", remove_noise="no", remove_dc="no", unit_area="yes", doublebw="no")
PL_DEFAULTS = dict(pl_window="5", pl_beam="1e-10", pl_pbeam="1e-10", pl_pip="1.0", pl_weight="3.0")


def _logs(p, logbase=1.0001):
    return (-(1 << 31) >> 2 if p <= 0 else int(math.log(p) * (1.0 / math.log(logbase)))) >> 10


class Decoder:
    def __init__(self, hmm, dict_file, lm_file, max_utts=64, max_frames=1 << 16, device=0, **config):
        cfg = {k: str(v) for k, v in config.items()}
        self.pm = PackedModel.from_dir(hmm, **{k: v for k, v in cfg.items() if k in ("varfloor", "tmatfloor", "mixwfloor", "topn", "ds", "aw")})
        # front end: the reference's own defaults (config_macro.h), overlaid by the model's feat.params, then by the
        # caller -- like ps_init; whatever the device front end does not implement is refused, not ignored
        from .s3io import read_feat_params
        fp = dict(FE_REFERENCE_DEFAULTS)
        fp.update(read_feat_params(os.path.join(hmm, "feat.params")))
        fp.update({k: v for k, v in cfg.items() if k in FE_REFERENCE_DEFAULTS})
        yes = ("yes", "1", "true", "True")
        if fp["cmn"] == "current":                       # the reference's alias for batch (cmn.c: cmn_type_str)
            fp["cmn"] = "batch"
        unsupported = []
        if fp["feat"] != "1s_c_d_dd": unsupported.append("-feat " + fp["feat"])
        if fp["cmn"] != "batch": unsupported.append("-cmn " + fp["cmn"])
        if fp["agc"] != "none": unsupported.append("-agc " + fp["agc"])
        if fp["varnorm"] in yes: unsupported.append("-varnorm yes")
        if fp["lda"]: unsupported.append("-lda")
        if fp["svspec"] not in ("", "0-12/13-25/26-38"): unsupported.append("-svspec " + fp["svspec"])
        if fp["dither"] in yes: unsupported.append("-dither yes")
        if fp["round_filters"] not in yes: unsupported.append("-round_filters no")
        if int(fp["ncep"]) != 13: unsupported.append("-ncep " + fp["ncep"])
        if int(fp["frate"]) != 100: unsupported.append("-frate " + fp["frate"])
        if int(fp["nfft"]) != 0: unsupported.append("-nfft " + fp["nfft"])
        if fp["cmninit"] not in ("", FE_REFERENCE_DEFAULTS["cmninit"]) and fp["cmn"] != "batch": unsupported.append("-cmninit")
        if unsupported:
            raise NotImplementedError("front end settings the device front end does not implement: " + ", ".join(unsupported))
        self.fe = api.FrontEnd(make_fe_desc(samprate=float(fp["samprate"]), wlen=float(fp["wlen"]), nfilt=int(fp["nfilt"]),
                                            lowerf=float(fp["lowerf"]), upperf=float(fp["upperf"]), alpha=float(fp["alpha"]),
                                            transform=fp["transform"], lifter=int(fp["lifter"]), remove_noise=fp["remove_noise"] in yes,
                                            remove_dc=fp["remove_dc"] in yes, unit_area=fp["unit_area"] in yes,
                                            doublebw=fp["doublebw"] in yes), device)
        search_cfg = {k: v for k, v in cfg.items() if k in lextree.DEFAULTS}
        self.search = lextree.ngram_search_from_files(hmm, dict_file, lm_file, **search_cfg)
        self.model = api.Model(self.pm, device)
        self.batch = api.Batch(self.model, max_utts, max_frames)
        self.ctx = api.HmmContext(self.pm.tp, self.pm.sseq, self.pm.n_sen, device=device)
        self.device = device
        pl = dict(PL_DEFAULTS)
        pl.update({k: v for k, v in cfg.items() if k in PL_DEFAULTS})
        self.pl_window = int(pl["pl_window"])
        n_ci = self.pm.n_ciphone
        self.phoneloop = api.PhoneLoop(self.ctx, self.pm.phone_ssid[:n_ci], self.pm.phone_tmat[:n_ci], self.pl_window,
                                       _logs(float(pl["pl_beam"])), _logs(float(pl["pl_pbeam"])), _logs(float(pl["pl_pip"])),
                                       float(pl["pl_weight"]))
        self.second_pass = search_cfg.get("fwdflat", "yes") in ("yes", "1", "true", "True")
        self.bp_cap = int(cfg.get("latsize", 5000))

    def decode_raw_batch(self, utterances):
        """utterances: int16 arrays, each a whole utterance.  Returns one dict per utterance: hyp (the words, fillers
        and <s> / </s> left out), score, seg [n][7] = entry, wid, sf, ef, path score, ascr, lscr, and words()."""
        import torch
        g = self.search
        info = g["info"]
        off = api.FrontEnd.sample_offsets([len(u) for u in utterances])
        pcm = np.concatenate([np.ascontiguousarray(u, np.int16) for u in utterances]) if utterances else np.zeros(0, np.int16)
        frame_off, best, pen = self.batch.decode_pcm_host(self.fe, self.phoneloop, pcm, off)
        d_scr = self.batch.senscr_device_ptr()
        d_pen = (torch.from_numpy(np.ascontiguousarray(pen, np.int32)).to(torch.device("cuda", self.device))
                 if self.pl_window > 0 and len(pen) else None)
        pen_ptr, win = (d_pen.data_ptr(), self.pl_window) if d_pen is not None else (None, 0)
        # -latsize is an initial size in the reference (bp_table / bscore_stack double on demand,
        # ngram_search.c:326-339): a table that fills up is retried with doubled capacities
        # ... and streams start from a size that fits them (the 72 k-word en-us LM writes about eleven entries per frame
        # and twenty scores per entry on the reference's test data): a retry repeats the whole search, the reference's
        # realloc does not
        longest = int(np.diff(frame_off).max()) if len(frame_off) > 1 else 0
        cap = max(self.bp_cap, 12 * longest + 1000)
        for _ in range(6):
            try:
                if self.second_pass:
                    tabs, _ = self.ctx.ngram_two_pass(d_scr, frame_off, info, g["model"], g["ci_tmat"], g["ci_ssid"], cap, 24 * cap,
                                                      pen_ptr, win, first_cap=cap, first_bss_cap=24 * cap, lm_arrays=g["lm_arrays"])
                else:
                    tabs = self.ctx.ngram_fwdtree(d_scr, frame_off, info, g["model"], g["ci_tmat"], cap, 24 * cap, pen_ptr, win,
                                                  lm_arrays=g["lm_arrays"])
                break
            except api.PsbError as e:
                if "overflow" not in str(e):
                    raise
                cap *= 2
        else:
            raise RuntimeError("backpointer table still overflows at %d entries per utterance" % cap)
        out = []
        words, base, fs, fe_ = g["words"], g["base"], int(info[22]), int(info[23])
        for u, (bp, bss, idx) in enumerate(tabs):
            T = int(frame_off[u + 1] - frame_off[u])
            entry, score, _ = api.ngram_hyp(bp, idx, T, int(info[20]))
            seg = api.ngram_segments(info, g["model"], bp, bss, entry, lm_arrays=g["lm_arrays"], second_pass=self.second_pass)
            real = [words[base[w]] for w in seg[:, 1] if not (fs <= int(base[w]) <= fe_)]        # dict_real_word + dict_basestr
            out.append(dict(hyp=" ".join(real), score=score, seg=seg, words=[words[w] for w in seg[:, 1]], n_frames=T))
        return out

    def close(self):
        for o in (self.phoneloop, self.batch, self.ctx, self.model, self.fe):
            o.close()
