"""Embeds the kernel headers into the shared object for run-time compilation of user right-hand sides (nvrtc_rhs.cu).
Usage: gen_rtc_headers.py OUT.inc name=path [name=path ...] — `name` is the string the headers use in their #include."""
import sys

out, pairs = sys.argv[1], [a.split("=", 1) for a in sys.argv[2:]]
with open(out, "w") as f:
    for name, path in pairs:
        text = open(path).read()
        assert ')VOHDR"' not in text
        f.write('{"%s",\n' % name)
        # a string literal may not exceed 64 KiB on every host compiler: emit the file in chunks that the compiler concatenates
        for k in range(0, len(text), 8000):
            f.write('R"VOHDR(' + text[k:k + 8000] + ')VOHDR"\n')
        f.write("},\n")
