#!/usr/bin/env python
"""SASS excerpts of libsdn_b200.so for profiles/: per kernel (name regex), the lines whose mnemonic matches a pattern.
   python scripts/sass_excerpts.py 'k_conv_wgradILi128' 'UTMALDG|UTCHMMA|UTCBAR|LDTM|REDG|SYNCS' [max_lines]"""
import re, subprocess, sys
so = '3d-sdn_b200/sdn_b200/libsdn_b200.so'
name, pat = sys.argv[1], re.compile(sys.argv[2])
limit = int(sys.argv[3]) if len(sys.argv) > 3 else 3
out = subprocess.run(['cuobjdump', '-sass', so], capture_output=True, text=True).stdout.splitlines()
cur, body = None, {}
for l in out:
    m = re.search(r'Function : (\S+)', l)
    if m:
        cur = m.group(1)
        continue
    if cur and re.search(name, cur) and re.match(r'\s+/\*[0-9a-f]{4,}\*/', l):
        body.setdefault(cur, []).append(l)
for fn, lines in body.items():
    hits = [re.sub(r'\s*/\* 0x[0-9a-f]+ \*/\s*$', '', x) for x in lines if pat.search(x)]
    counts = {}
    for x in lines:
        m = pat.search(x)
        if m:
            counts[m.group(0)] = counts.get(m.group(0), 0) + 1
    print('## %s  (%d SASS instructions; matches: %s)' % (fn[:110], len(lines), ', '.join('%s x%d' % kv for kv in sorted(counts.items()))))
    seen = {}
    for x in hits:                      # at most `limit` lines per mnemonic, in program order
        k = pat.search(x).group(0)
        seen[k] = seen.get(k, 0) + 1
        if seen[k] <= limit:
            print(x.rstrip())
    print()
