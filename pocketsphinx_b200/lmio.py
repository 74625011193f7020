"""Binary trie language models (`*.lm.bin`) as the arrays the n-gram search kernels score with.

The reference keeps a loaded LM as a bit-packed trie (lm/lm_trie.c, lm/bitarr.c, 16-bit quantised
probabilities lm/lm_trie_quant.c) and its maintainer-side binding unpacks that trie for the device
(integration/ps_search_cuda.c:cuda_ngram_export_lm).  A Python host has no trie to unpack: this module
reads the file itself (ngram_model_trie_read_bin, ngram_model_trie.c:372-443) and produces the same
int32 block -- header [10] | widmap | unigram prob / backoff / next | bigram word / prob / backoff / next |
trigram word / prob (psb_ngram_desc_t.lm_arrays, scoring in csrc/psb_lm_core.h).
tests/test_lmio.py compares the block with the binding's, word for word, on the reference's own models.
"""
import math
import struct

import numpy as np

TRIE_HDR = b"Trie Language Model"


def _required_bits(v):
    """bitarr_required_bits (lm/bitarr.c)."""
    return int(v).bit_length()


def _unpack(mem, n, total_bits, offset_bits, width):
    """bitarr_read_int25 for entries 0..n-1 of a packed array: the 32-bit little-endian word at the
    field's byte, shifted by the bit remainder, masked to `width` bits."""
    if n == 0:
        return np.zeros(0, np.uint32)
    bit = np.arange(n, dtype=np.int64) * total_bits + offset_bits
    byte = bit >> 3
    b = mem.astype(np.uint32)
    word = b[byte] | (b[byte + 1] << 8) | (b[byte + 2] << 16) | (b[byte + 3] << 24)
    return (word >> (bit & 7).astype(np.uint32)) & np.uint32((1 << width) - 1)


def read_lm_bin(path):
    """-> dict(order, counts, words, uni_prob, uni_bo, uni_next [V+1], and for order >= 2 bg_word,
    bg_prob, bg_bo, bg_next [n2+1], for order 3 tg_word, tg_prob); probabilities are the float32 values
    the trie holds (log domain of the file's log base, language weight not applied)."""
    with open(path, "rb") as f:
        raw = f.read()
    if not raw.startswith(TRIE_HDR):
        raise ValueError("%s: not a binary trie LM" % path)
    pos = len(TRIE_HDR)
    order = raw[pos]
    pos += 1
    if order < 1 or order > 3:
        raise NotImplementedError("%s: %d-gram model (orders 1..3 are scored on the device)" % (path, order))
    counts = struct.unpack_from("<%dI" % order, raw, pos)
    pos += 4 * order
    V = counts[0]
    quant = None
    if order > 1:                                            # lm_trie_quant_read_bin: a dummy word, then the bin centres
        nvalues = (order - 2) * (2 << 16) + (1 << 16)
        quant = np.frombuffer(raw, "<f4", nvalues, pos + 4)
        pos += 4 + 4 * nvalues
    ug = np.frombuffer(raw, np.dtype([("prob", "<f4"), ("bo", "<f4"), ("next", "<u4")]), V + 1, pos)
    pos += 12 * (V + 1)
    out = dict(order=order, counts=list(counts), uni_prob=ug["prob"][:V].copy(), uni_bo=ug["bo"][:V].copy(),
               uni_next=ug["next"].astype(np.int64))
    word_bits = _required_bits(V)

    def size(entries, remaining):                            # base_size (lm_trie.c:59-67)
        return ((1 + entries) * (word_bits + remaining) + 7) // 8 + 8
    if order == 2:
        n2 = counts[1]
        total = word_bits + 16
        mem = np.frombuffer(raw, np.uint8, size(n2, 16), pos)
        pos += len(mem)
        out["bg_word"] = _unpack(mem, n2, total, 0, word_bits).astype(np.int64)
        out["bg_prob"] = quant[_unpack(mem, n2, total, word_bits, 16)]           # the longest order's table
        out["bg_bo"] = np.zeros(n2, np.float32)
        out["bg_next"] = np.zeros(n2 + 1, np.int64)
    elif order == 3:
        n2, n3 = counts[1], counts[2]
        next_bits = _required_bits(n3)
        total = word_bits + 32 + next_bits                   # middle: word | backoff bin | probability bin | next
        mem = np.frombuffer(raw, np.uint8, size(n2, 32 + next_bits), pos)
        pos += len(mem)
        out["bg_word"] = _unpack(mem, n2, total, 0, word_bits).astype(np.int64)
        out["bg_bo"] = quant[(1 << 16) + _unpack(mem, n2, total, word_bits, 16)]
        out["bg_prob"] = quant[_unpack(mem, n2, total, word_bits + 16, 16)]
        out["bg_next"] = _unpack(mem, n2 + 1, total, word_bits + 32, next_bits).astype(np.int64)
        total = word_bits + 16
        mem = np.frombuffer(raw, np.uint8, size(n3, 16), pos)
        pos += len(mem)
        out["tg_word"] = _unpack(mem, n3, total, 0, word_bits).astype(np.int64)
        out["tg_prob"] = quant[(2 << 16) + _unpack(mem, n3, total, word_bits, 16)]
    k = struct.unpack_from("<i", raw, pos)[0]
    pos += 4
    if k < 0 or pos + k > len(raw):
        raise ValueError("%s: word strings leave the file" % path)
    words = raw[pos:pos + k].split(b"\0")[:-1]
    if len(words) != V:
        raise ValueError("%s: %d word strings for %d unigrams" % (path, len(words), V))
    out["words"] = [w.decode("latin-1") for w in words]
    return out


def lm_arrays(lm, dict_words, lw=6.5, wip=0.65, logbase=1.0001):
    """The int32 block of psb_ngram_desc_t.lm_arrays for a dictionary whose word strings, in id order,
    are dict_words (alternate pronunciations "word(2)" are looked up as written, like
    ngram_model_set_map_words does, ngram_model_set.c:640-669; words the LM does not know map to <UNK>
    or -1).  lw / wip: -lw / -wip of the first pass (ngram_model_apply_weights)."""
    order, V = lm["order"], lm["counts"][0]
    n2 = lm["counts"][1] if order >= 2 else 0
    n3 = lm["counts"][2] if order >= 3 else 0
    wid = {}
    for i, w in enumerate(lm["words"]):
        wid.setdefault(w, i)                                 # duplicates: the first one wins (hash_table_enter)
    unk = wid.get("<UNK>", -1)
    widmap = np.array([wid.get(w, unk) for w in dict_words], np.int32)
    log_wip = int(math.log(wip) * (1.0 / math.log(logbase))) if wip > 0 else -(1 << 31) >> 2      # logmath_log, shift 0
    hdr = np.zeros(10, np.int32)
    hdr[0:4] = (order, V, n2, n3)
    hdr[4] = np.array([lw], np.float32).view(np.int32)[0]
    hdr[5], hdr[6], hdr[7] = log_wip, -(1 << 31) >> 2, len(dict_words)
    hdr[8] = V if order >= 2 else 0                          # max_vocab of the bigram array (lm_trie_alloc_ngram: counts[0])
    hdr[9] = V if order == 3 else 0

    def f2i(a):
        return np.ascontiguousarray(a, np.float32).view(np.int32)
    parts = [hdr, widmap, f2i(lm["uni_prob"]), f2i(lm["uni_bo"]), lm["uni_next"].astype(np.int32)]
    if order >= 2:
        parts += [lm["bg_word"].astype(np.int32), f2i(lm["bg_prob"]), f2i(lm["bg_bo"]), lm["bg_next"].astype(np.int32)]
    else:
        parts += [np.zeros(1, np.int32)]                     # bg_next [n2 + 1] with n2 = 0
    if order == 3:
        parts += [lm["tg_word"].astype(np.int32), f2i(lm["tg_prob"])]
    return np.concatenate(parts).astype(np.int32)


def read_dict(dictfile, fillerfile=None, ciphones=None):
    """The pronunciation dictionary in the reference's word-id order (dict_init / dict_read / dict_add_word,
    dict.c:77-145, 146-215, 258-395): the main file's entries, then the filler file's (an acoustic model's
    `noisedict`), then <s>, </s>, <sil> if nothing defined them.  Lines starting with "##" or ";;" are
    comments; entries without a pronunciation, with a phone the model lacks (ciphones: the mdef's CI phone
    names), duplicates and alternates "word(2)" whose base word is missing are dropped like the reference
    drops them.  Returns (words, prons, basewid, filler_start): prons as lists of phone names."""
    words, prons, base, index = [], [], [], {}
    known = None if ciphones is None else set(ciphones)

    def add(word, phones):
        b = len(words)
        if word.endswith(")"):
            i = word.rfind("(", 1, len(word) - 1)
            if i > 0:
                if word[:i] not in index:
                    return False
                b = index[word[:i]]
        if word in index:
            return False
        index[word] = len(words)
        words.append(word); prons.append(phones); base.append(b)
        return True

    def read(path):
        with open(path, "r", encoding="latin-1") as f:
            for line in f:
                if line.startswith("##") or line.startswith(";;"):
                    continue
                t = line.split()
                if len(t) < 2:
                    continue
                if known is not None and any(p not in known for p in t[1:]):
                    continue
                add(t[0], t[1:])
    read(dictfile)
    for w in ("<s>", "</s>", "<sil>"):
        if w in index:
            raise ValueError("%s: remove '%s' from the dictionary" % (dictfile, w))
    filler_start = len(words)
    if fillerfile:
        read(fillerfile)
    for w in ("<s>", "</s>", "<sil>"):
        if w not in index:
            add(w, ["SIL"])
    return words, prons, np.array(base, np.int32), filler_start
