"""One-off source transformation used to make kernels batch-aware (batch.cuh): rewrites
    __global__ void [__launch_bounds__(B)] name(params) {      ->  GINGR_KERNEL((B), name, params) {
    name<<<grid, block[, smem[, stream]]>>>(args);              ->  GINGR_LAUNCH(ctx, name, grid, block, smem, stream, args);
for the kernel names given on the command line, in the files of gingr_b200/csrc.  Template kernels are converted by hand.
usage: python tools/batchify_kernels.py name [name ...]"""
import os
import re
import sys

CSRC = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gingr_b200", "csrc")


def match_paren(s, i):
    """s[i] == '(' -> index of the matching ')' (comments skipped)"""
    depth = 0
    k = i
    while k < len(s):
        if s.startswith("/*", k):
            k = s.index("*/", k) + 2
            continue
        if s.startswith("//", k):
            k = s.index("\n", k)
            continue
        c = s[k]
        if c == "(":
            depth += 1
        elif c == ")":
            depth -= 1
            if depth == 0:
                return k
        k += 1
    raise ValueError("unbalanced")


def split_top(s):
    out, depth, cur = [], 0, ""
    for c in s:
        if c in "([{":
            depth += 1
        elif c in ")]}":
            depth -= 1
        if c == "," and depth == 0:
            out.append(cur.strip())
            cur = ""
        else:
            cur += c
    if cur.strip():
        out.append(cur.strip())
    return out


def convert(text, name):
    ndef = nl = 0
    # definition
    m = re.search(r"__global__ void (?:__launch_bounds__\(([^\n]*?)\)\s+)?" + re.escape(name) + r"\(", text)
    if m:
        pre = text[:m.start()]
        if re.search(r"template\s*<[^;{}]*>\s*$", pre):
            print("  (template definition of", name, "left alone)")
        else:
            close = match_paren(text, m.end() - 1)
            params = text[m.end():close]
            bounds = m.group(1)
            head = (f"GINGR_KERNEL(({bounds}), {name}, " if bounds else f"GINGR_KERNEL_NB({name}, ") + params + ")"
            text = text[:m.start()] + head + text[close + 1:]
            ndef = 1
    # launches
    pos = 0
    pat = re.compile(r"\b" + re.escape(name) + r"<<<")
    while True:
        m = pat.search(text, pos)
        if not m:
            break
        cfg_end = text.index(">>>", m.end())
        cfg = split_top(text[m.end():cfg_end])
        assert 2 <= len(cfg) <= 4, (name, cfg)
        while len(cfg) < 3:
            cfg.append("0")
        if len(cfg) < 4:
            cfg.append("0")
        a0 = cfg_end + 3
        assert text[a0] == "(", (name, text[a0:a0 + 20])
        a1 = match_paren(text, a0)
        args = text[a0 + 1:a1]
        rep = f"GINGR_LAUNCH(ctx, {name}, {cfg[0]}, {cfg[1]}, {cfg[2]}, {cfg[3]}, {args})"
        text = text[:m.start()] + rep + text[a1 + 1:]
        pos = m.start() + len(rep)
        nl += 1
    return text, ndef, nl


def main():
    names = sys.argv[1:]
    for f in sorted(os.listdir(CSRC)):
        if not f.endswith((".cu", ".cuh")) or f == "batch.cuh":
            continue
        p = os.path.join(CSRC, f)
        text = open(p).read()
        orig = text
        for n in names:
            text, d, l = convert(text, n)
            if d or l:
                print(f"{f}: {n}: {d} definition, {l} launches")
        if text != orig:
            open(p, "w").write(text)


if __name__ == "__main__":
    main()
