#!/usr/bin/env python
"""Static SASS instruction count per source line of one kernel (nvdisasm -g output).  usage: sass_lines.py <file.sass> <substr> [top]"""
import re, collections, sys
lines = open(sys.argv[1]).read().split('\n')
key = sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 30
start = [i for i, l in enumerate(lines) if '.text.' in l and key in l and '------' in l][0]
end = [i for i, l in enumerate(lines) if i > start + 5 and '------' in l and '.text.' in l]
end = end[0] if end else len(lines)
cur, cnt = None, collections.Counter()
for l in lines[start:end]:
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m:
        cur = (m.group(1), int(m.group(2)))
        continue
    if re.search(r'/\*[0-9a-f]{4,6}\*/\s', l) and cur:
        cnt[cur] += 1
print('static instructions', sum(cnt.values()))
cache = {}
for (f, ln), c in sorted(cnt.items(), key=lambda kv: -kv[1])[:top]:
    try:
        if f not in cache:
            cache[f] = open(f).read().split('\n')
        t = cache[f][ln - 1].strip()[:100]
    except Exception:
        t = ''
    print('%4d  %-22s:%-4d %s' % (c, f.split('/')[-1], ln, t))
