#!/usr/bin/env python
"""Query-log generator after the reference's tools/gen_synthetic_log.py.

Input: a "term df" list (one per line). Terms are grouped by document frequency the way
gen_synthetic_log.py:21-28,82-88 does: exponent int(log10(df)) 0-3 -> "low", 4-6 -> "high"
(i.e. high = df >= 10^4); --high-df moves that boundary for small test corpora.

Workloads (one query per line, space-separated terms — query_pool.h:251-311 format):
  single_low / single_high : uniform sample WITH replacement from the group (:169-189)
  two_term                 : UNIQUE queries; each term's group drawn uniformly from
                             {low, high}, term uniform in group, t2 != t1, the two terms
                             sorted lexicographically (:191-211)
  multi_term (ours, config 3): 3-5 terms, at least one high and one low, unique terms
  mix_aol (ours)           : term-count mix after data/AOL_QueryLog_analysis/stat.txt:8-11
Deterministic for a given --seed (random.Random, not the global generator).
"""
import argparse
import random


def load_groups(path, high_df):
    groups = {"low": [], "high": []}
    with open(path) as f:
        for line in f:
            items = line.split()
            if len(items) != 2:
                continue
            df = int(items[1])
            groups["high" if df >= high_df else "low"].append(items[0])
    # file order depends on hash-map iteration in the indexer; sort for reproducibility
    groups["low"].sort()
    groups["high"].sort()
    return groups


def single_term(rng, groups, group, n):
    g = groups[group]
    return [g[rng.randint(0, len(g) - 1)] for _ in range(n)]


def two_term_fixed(rng, groups, n, g1name, g2name):
    """Diagnostic variants of two_term with the groups pinned (hh / lh / ll)."""
    queries, seen = [], set()
    limit = n * 50
    g1, g2 = groups[g1name], groups[g2name]
    while len(queries) < n and limit > 0:
        limit -= 1
        t1 = g1[rng.randint(0, len(g1) - 1)]
        t2 = g2[rng.randint(0, len(g2) - 1)]
        if t2 == t1:
            continue
        q = " ".join(sorted([t1, t2]))
        if q not in seen:
            seen.add(q)
            queries.append(q)
    return queries


def phrase_queries(rng, groups, n, m=2):
    """Quoted m-term phrase queries (gen_synthetic_log.py:254-265 wraps sampled phrases in double
    quotes). Terms are drawn from the high-df group so that phrases do occur in a synthetic corpus."""
    g = groups["high"] or groups["low"]
    out = []
    for _ in range(n):
        out.append('"' + " ".join(g[rng.randint(0, len(g) - 1)] for _ in range(m)) + '"')
    return out


def two_term(rng, groups, n):
    names = [g for g in ("low", "high") if groups[g]]
    queries, seen = [], set()
    limit = n * 50
    while len(queries) < n and limit > 0:
        limit -= 1
        g1 = groups[names[rng.randint(0, len(names) - 1)]]
        g2 = groups[names[rng.randint(0, len(names) - 1)]]
        t1 = g1[rng.randint(0, len(g1) - 1)]
        t2 = g2[rng.randint(0, len(g2) - 1)]
        if t2 == t1:
            continue
        q = " ".join(sorted([t1, t2]))
        if q not in seen:
            seen.add(q)
            queries.append(q)
    return queries


def multi_term(rng, groups, n, lo=3, hi=5):
    out = []
    for _ in range(n):
        m = rng.randint(lo, hi)
        n_high = rng.randint(1, m - 1)
        terms = set()
        guard = 0
        while len(terms) < n_high and guard < 1000:
            terms.add(groups["high"][rng.randint(0, len(groups["high"]) - 1)])
            guard += 1
        while len(terms) < m and guard < 2000:
            terms.add(groups["low"][rng.randint(0, len(groups["low"]) - 1)])
            guard += 1
        t = sorted(terms)   # set order depends on the per-process hash seed
        rng.shuffle(t)
        out.append(" ".join(t))
    return out


def mix_aol(rng, groups, n):
    # 1..4+ term shares from the AOL log statistics (renormalised)
    shares = [(1, 36.8), (2, 25.2), (3, 17.3), (4, 10.0), (5, 10.7)]
    tot = sum(s for _, s in shares)
    out = []
    allterms = groups["low"] + groups["high"]
    for _ in range(n):
        r = rng.random() * tot
        m = 1
        for cnt, s in shares:
            if r < s:
                m = cnt
                break
            r -= s
        if m == 1:
            g = "high" if (groups["high"] and rng.random() < 0.5) else "low"
            out.append(single_term(rng, groups, g, 1)[0])
        elif m == 2:
            out.extend(two_term(rng, groups, 1) or [allterms[0]])
        else:
            out.extend(multi_term(rng, groups, 1, m, m))
    return out


def generate(kind, groups, n, seed):
    rng = random.Random(seed)
    if kind == "single_low":
        return single_term(rng, groups, "low", n)
    if kind == "single_high":
        return single_term(rng, groups, "high", n)
    if kind == "two_term":
        return two_term(rng, groups, n)
    if kind == "two_term_hh":
        return two_term_fixed(rng, groups, n, "high", "high")
    if kind == "two_term_lh":
        return two_term_fixed(rng, groups, n, "low", "high")
    if kind == "two_term_ll":
        return two_term_fixed(rng, groups, n, "low", "low")
    if kind == "phrase2":
        return phrase_queries(rng, groups, n, 2)
    if kind == "phrase3":
        return phrase_queries(rng, groups, n, 3)
    if kind == "multi_term":
        return multi_term(rng, groups, n)
    if kind == "mix_aol":
        return mix_aol(rng, groups, n)
    raise ValueError(kind)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--terms", required=True, help="file with 'term df' per line")
    ap.add_argument("--kind", required=True,
                    choices=["single_low", "single_high", "two_term", "two_term_hh", "two_term_lh",
                             "two_term_ll", "phrase2", "phrase3", "multi_term", "mix_aol"])
    ap.add_argument("--n", type=int, required=True)
    ap.add_argument("--high-df", type=int, default=10000)
    ap.add_argument("--seed", type=int, default=1)
    ap.add_argument("--out", required=True)
    a = ap.parse_args()
    groups = load_groups(a.terms, a.high_df)
    qs = generate(a.kind, groups, a.n, a.seed)
    with open(a.out, "w") as f:
        for q in qs:
            f.write(q + "\n")


if __name__ == "__main__":
    main()
