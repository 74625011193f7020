"""Batch n-gram decoding from audio to words on the device, configured like the reference's ps_decoder_t
(-hmm, -dict, -lm and the search settings by their reference names): front end -> senone scores -> phone
loop -> first pass -> second pass on the GPU, hypothesis and segments read from the returned tables
(ps_decode_raw + ps_get_hyp / ps_seg_iter with -bestpath no, pocketsphinx.c:1073-1345).  Everything the
reference loads from files is read here by the package itself (s3io, lmio, dict2pid, lextree).

STATUS: the loaders and the search description are pinned against the reference on the CPU; the three
stages in front of the search are GPU-verified; the search kernels themselves have not run on hardware yet
(DESIGN.md 4.10-4.12), so neither has this class -- tests/test_gpu_zz_decoder.py is gated like theirs.
"""
import math
import os

import numpy as np

from . import api, lextree
from .fe_tables import make_fe_desc
from .model import PackedModel

PL_DEFAULTS = dict(pl_window="5", pl_beam="1e-10", pl_pbeam="1e-10", pl_pip="1.0", pl_weight="3.0")


def _logs(p, logbase=1.0001):
    return (-(1 << 31) >> 2 if p <= 0 else int(math.log(p) * (1.0 / math.log(logbase)))) >> 10


class Decoder:
    def __init__(self, hmm, dict_file, lm_file, max_utts=64, max_frames=1 << 16, device=0, **config):
        cfg = {k: str(v) for k, v in config.items()}
        self.pm = PackedModel.from_dir(hmm, **{k: v for k, v in cfg.items() if k in ("varfloor", "tmatfloor", "mixwfloor", "topn", "ds", "aw")})
        fp = {}
        from .s3io import read_feat_params
        fp.update(read_feat_params(os.path.join(hmm, "feat.params")))
        if fp.get("feat", "1s_c_d_dd") != "1s_c_d_dd" or fp.get("cmn", "batch") != "batch":
            raise NotImplementedError("front end: -feat %s / -cmn %s (the device front end covers 1s_c_d_dd with batch CMN)" % (
                fp.get("feat"), fp.get("cmn")))
        fe_kw = {}
        for k, name, conv in (("nfilt", "nfilt", int), ("lowerf", "lowerf", float), ("upperf", "upperf", float), ("lifter", "lifter", int),
                              ("transform", "transform", str), ("samprate", "samprate", float), ("wlen", "wlen", float)):
            if k in fp:
                fe_kw[name] = conv(fp[k])
        if "remove_noise" in fp:
            fe_kw["remove_noise"] = fp["remove_noise"] in ("yes", "1", "true")
        if "remove_dc" in fp:
            fe_kw["remove_dc"] = fp["remove_dc"] in ("yes", "1", "true")
        self.fe = api.FrontEnd(make_fe_desc(**fe_kw), device)
        search_cfg = {k: v for k, v in cfg.items() if k in lextree.DEFAULTS}
        self.search = lextree.ngram_search_from_files(hmm, dict_file, lm_file, **search_cfg)
        self.model = api.Model(self.pm, device)
        self.batch = api.Batch(self.model, max_utts, max_frames)
        self.ctx = api.HmmContext(self.pm.tp, self.pm.sseq, self.pm.n_sen)
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
        d_pen = torch.from_numpy(np.ascontiguousarray(pen, np.int32)).cuda() if self.pl_window > 0 and len(pen) else None
        pen_ptr, win = (d_pen.data_ptr(), self.pl_window) if d_pen is not None else (None, 0)
        if self.second_pass:
            tabs, _ = self.ctx.ngram_two_pass(d_scr, frame_off, info, g["model"], g["ci_tmat"], g["ci_ssid"], self.bp_cap, 20 * self.bp_cap,
                                              pen_ptr, win, first_cap=self.bp_cap, first_bss_cap=20 * self.bp_cap, lm_arrays=g["lm_arrays"])
        else:
            tabs = self.ctx.ngram_fwdtree(d_scr, frame_off, info, g["model"], g["ci_tmat"], self.bp_cap, 20 * self.bp_cap, pen_ptr, win,
                                          lm_arrays=g["lm_arrays"])
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
