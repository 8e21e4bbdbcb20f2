"""build helper: embeds bp_kernels.cuh into libbpcuda as a C string so NVRTC can #include it."""
import sys

src, dst = sys.argv[1], sys.argv[2]
text = open(src).read()
assert ')BPKERN"' not in text
chunks = [text[i:i + 8000] for i in range(0, len(text), 8000)]
with open(dst, "w") as f:
    f.write('// generated from bp_kernels.cuh by embed_header.py\nextern "C" const char bp_kernels_cuh_text[] =\n')
    for c in chunks:
        f.write('R"BPKERN(' + c + ')BPKERN"\n')
    f.write(";\n")
