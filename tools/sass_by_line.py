"""Static SASS instruction count per source line of a scene-specialised kernel:  python tools/sass_by_line.py PREFIX [top]
(PREFIX.cubin / PREFIX.cu from tools/jit_offline.py; uses `nvdisasm -g`)."""
import collections, re, subprocess, sys
prefix = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
out = subprocess.run(["nvdisasm", "-g", "-c", prefix + ".cubin"], capture_output=True, text=True).stdout
src = open(prefix + ".cu").read().splitlines()
cur, cnt, infun = None, collections.Counter(), False
for line in out.splitlines():
    if line.startswith(".text."):
        infun = "k_bounce_jit" in line
    m = re.search(r'//## File "[^"]+", line (\d+)', line)
    if m:
        cur = int(m.group(1))
        continue
    if infun and re.match(r"\s+/\*[0-9a-f]{4,}\*/", line):
        cnt[cur] += 1
print("k_bounce_jit: static instructions", sum(cnt.values()))
for ln, c in cnt.most_common(top):
    print("%5d %4d  %s" % (ln or 0, c, src[ln - 1].strip()[:120] if ln and ln <= len(src) else ""))
