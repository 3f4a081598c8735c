#!/usr/bin/env python
"""Seeded synthetic Zipf corpus in the reference's linedoc WITH_POSITIONS format.

Format (reference: src/qq_mem/src/testdata/iter_test_3_docs:1-4, engine_loader.h:84-97):
  header  FIELDS_HEADER_INDICATOR###\\tdoctitle\\tbody\\ttokenized\\toffsets\\tpositions
  per doc title \\t body \\t unique terms (space-sep) \\t per-term "s,e;s,e;." \\t per-term "p;p;."
Doc ids are assigned 0..N-1 in file order by the indexer (flash_engine_dumper.h:714-721).
The body is the token sequence joined by single spaces, so the reference's BodyLength()
(token count) equals the sampled document length.

Corpus model: term rank ~ Zipf(s) over V terms named t<rank>; document
length ~ clip(lognormal(mu, sigma), lo, hi).
Used to make small corpora that the REFERENCE indexer turns into golden fixtures
(tests/golden/make_golden.py); large corpora come from the native generator instead.
"""
import argparse

import numpy as np


def zipf_cdf(vocab, s):
    w = 1.0 / np.power(np.arange(1, vocab + 1, dtype=np.float64), s)
    c = np.cumsum(w)
    return c / c[-1]


def doc_tokens(rng, cdf, length):
    return np.searchsorted(cdf, rng.random(length), side="right")


def format_doc(doc_id, ranks, prefix="t"):
    toks = [f"{prefix}{r}" for r in ranks]
    body = " ".join(toks)
    uniq, offs, poss = [], {}, {}
    cur = 0
    for pos, t in enumerate(toks):
        if t not in offs:
            uniq.append(t)
            offs[t] = []
            poss[t] = []
        offs[t].append((cur, cur + len(t)))
        poss[t].append(pos)
        cur += len(t) + 1
    off_s = "".join("".join(f"{s},{e};" for s, e in offs[t]) + "." for t in uniq)
    pos_s = "".join("".join(f"{p};" for p in poss[t]) + "." for t in uniq)
    return f"doc_{doc_id}\t{body}\t{' '.join(uniq)}\t{off_s}\t{pos_s}\n"


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--docs", type=int, required=True)
    ap.add_argument("--vocab", type=int, required=True)
    ap.add_argument("--zipf-s", type=float, default=1.0)
    ap.add_argument("--mu", type=float, default=4.3)
    ap.add_argument("--sigma", type=float, default=0.6)
    ap.add_argument("--min-len", type=int, default=5)
    ap.add_argument("--max-len", type=int, default=2000)
    ap.add_argument("--seed", type=int, default=1)
    ap.add_argument("--out", required=True)
    a = ap.parse_args()

    rng = np.random.default_rng(a.seed)
    cdf = zipf_cdf(a.vocab, a.zipf_s)
    with open(a.out, "w") as f:
        f.write("FIELDS_HEADER_INDICATOR###\tdoctitle\tbody\ttokenized\toffsets\tpositions\n")
        for d in range(a.docs):
            n = int(np.clip(rng.lognormal(a.mu, a.sigma), a.min_len, a.max_len))
            f.write(format_doc(d, doc_tokens(rng, cdf, n)))


if __name__ == "__main__":
    main()
