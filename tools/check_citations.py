#!/usr/bin/env python
"""Check that every `src/...rs:LINE[-LINE]` citation in the headers, sources and docs points inside an existing
file of the reference checkout (run where /root/reference is mounted; not part of the test-suite)."""
import os
import re
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = sys.argv[1] if len(sys.argv) > 1 else "/root/reference"
PAT = re.compile(r"((?:src|tests|benches)/[A-Za-z0-9_/]+\.rs):(\d+)(?:-(\d+))?((?:,:\d+(?:-\d+)?)*)")
files = []
for d, _, fs in os.walk(ROOT):
    if any(p in d for p in (".git", "gpurun_out", "__pycache__", ".pytest_cache")):
        continue
    files += [os.path.join(d, f) for f in fs if f.endswith((".h", ".hpp", ".cu", ".cuh", ".cpp", ".py", ".md", ".rs"))]
lens, bad, n = {}, [], 0
for f in files:
    if os.path.basename(f) in ("SURVEY.md", "PAPERS.md", "SNIPPETS.md"):
        continue
    for m in PAT.finditer(open(f, errors="ignore").read()):
        path = os.path.join(REF, m.group(1))
        if path not in lens:
            lens[path] = sum(1 for _ in open(path, errors="ignore")) if os.path.exists(path) else -1
        spans = [(m.group(2), m.group(3))] + re.findall(r",:(\d+)(?:-(\d+))?", m.group(4))
        for a, b in spans:
            n += 1
            hi = int(b) if b else int(a)
            if lens[path] < 0 or hi > lens[path] or int(a) > hi:
                bad.append((os.path.relpath(f, ROOT), m.group(1), a, b, lens[path]))
print(f"{n} citations checked, {len(bad)} invalid")
for b in bad[:40]:
    print("  ", b)
