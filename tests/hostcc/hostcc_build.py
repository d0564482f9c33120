"""TEST INFRASTRUCTURE ONLY: builds chord_b200/csrc/*.cu for the HOST (g++, stand-in headers of this directory) into a
scratch shared library with the same C ABI, so that the parity cases can execute the kernels' source text without a GPU.

What is rewritten on the way (text substitution on a copy, the product sources are not touched):
  * ``kernel<<<grid, block, 0, stream>>>(args);`` -> ``hostcc::launch(grid, block, [=] { kernel(args); });``
  * the bodies of the few helpers that are inline PTX (mbarrier / TMA bulk copy / cp.async / RED with an L2 policy / volatile
    loads / the MUFU-based reciprocal sequences) -> what the instruction does, synchronously; the approximate reciprocal
    sequences become the IEEE division they are proven equal to on the GPU (chord_selftest_vertex_records).
See cuda_runtime.h here for what a host run proves and what it does not."""
import os
import re
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
CSRC = os.path.join(ROOT, 'chord_b200', 'csrc')
SOURCES = ['ctx.cu', 'cull.cu', 'raster.cu', 'hzb.cu', 'resolve.cu', 'shadow.cu', 'renderer.cu']

# helper name -> host body (the signature stays the product's)
PTX_HELPERS = {
    'loads_issued': '{ (void)a; (void)b; (void)c; (void)d; }',
    'ld_volatile_u64': '{ hostcc::yield(); return *(const volatile unsigned long long*)p; }',
    'st_volatile_u64': '{ *(volatile unsigned long long*)p = v; }',
    'smem_u32': '{ return (uint32_t)(uintptr_t)p; }',
    # mbarrier as a count of completed phases: phase k is complete once the count exceeds k
    'mbar_init': '{ (void)count; *bar = 0ull; }',
    'mbar_expect_tx': '{ (void)bar; (void)bytes; }',
    'l2_policy_evict_first': '{ return 0ull; }',
    'l2_policy_evict_last': '{ return 0ull; }',
    'tma_bulk_g2s': '{ (void)policy; memcpy(dst, src, bytes); *bar += 1ull; }',
    'mbar_wait': '{ while (((uint32_t)(*bar) & 1u) == parity) hostcc::yield(); }',
    'cp_async_4': '{ (void)policy; memcpy(smemDst, gmemSrc, 4); }',
    'cp_async_commit': '{ }',
    'cp_async_wait_all': '{ }',
    'vis_red': '{ (void)policy; z = fminf(fmaxf(z, 0.0f), 1.0f); const unsigned long long key = ((unsigned long long)__float_as_uint(z) << 32) | (unsigned long long)visId; if (key > *p) *p = key; }',
    'ld_work_stream': '{ (void)policy; return *p; }',
    'rcp_rn_midrange': '{ return 1.0f / x; }',
    'div3_rn': '{ qx = x / w; qy = y / w; qz = z / w; }',
}


def _replace_body(src, name, body):
    m = re.search(r'\b' + name + r'\s*\([^;{}]*?\)\s*\{', src)
    assert m, f'helper {name} not found'
    i = m.end() - 1
    depth, j = 0, i
    while True:
        c = src[j]
        if c == '{':
            depth += 1
        elif c == '}':
            depth -= 1
            if depth == 0:
                break
        j += 1
    return src[:i] + body + src[j + 1:]


def _split_top(s):
    out, depth, cur = [], 0, ''
    for c in s:
        if c in '([{':
            depth += 1
        elif c in ')]}':
            depth -= 1
        if c == ',' and depth == 0:
            out.append(cur.strip()); cur = ''
        else:
            cur += c
    out.append(cur.strip())
    return out


def translate(text, name):
    n_launch = text.count('<<<')

    def launch(m):
        cfg = _split_top(m.group(2))
        return f'hostcc::launch(dim3({cfg[0]}), dim3({cfg[1]}), [=] {{ {m.group(1)}({m.group(3)}); }});'
    text, n = re.subn(r'([A-Za-z_]\w*(?:<[^<>();]*>)?)<<<(.*?)>>>\((.*?)\);', launch, text, flags=re.S)
    assert n == n_launch and '<<<' not in text, f'{name}: {n} of {n_launch} launches translated'
    for h, body in PTX_HELPERS.items():
        if re.search(r'\b' + h + r'\s*\([^;{}]*?\)\s*\{', text):
            text = _replace_body(text, h, body)
    text = re.sub(r'^\s*asm volatile\("fence\.mbarrier_init[^\n]*\n', '\n', text, flags=re.M)
    assert 'asm' not in re.sub(r'//[^\n]*', '', text), f'{name}: inline PTX left after translation'
    return text


def build(outdir, opt='-O1', extra=()):
    """Translate + compile; returns the path of the scratch library (same exported C ABI as libchord_b200.so).
    ``extra``: more compiler / linker flags, e.g. ('-fsanitize=undefined', '-fno-sanitize-recover=undefined') to run the
    kernels under UBSan (signed overflow of the int32 edge functions, misaligned vector loads, shifts)."""
    outdir = str(outdir)
    inc = os.path.join(outdir, 'csrc')
    os.makedirs(inc, exist_ok=True)
    open(os.path.join(inc, 'common.cuh'), 'w').write(
        translate(open(os.path.join(CSRC, 'common.cuh')).read().replace('"../../include/chord_b200.h"', '<chord_b200.h>'), 'common.cuh'))
    cxx = '/usr/bin/g++' if os.path.exists('/usr/bin/g++') else 'g++'
    flags = [opt, '-std=c++17', '-fPIC', '-ffp-contract=off', '-mfma', '-mf16c', '-fno-fast-math', '-w',
             '-I', HERE, '-I', inc, '-I', os.path.join(ROOT, 'include')] + list(extra)
    objs, procs = [], []
    for s in SOURCES:
        cpp = os.path.join(inc, s.replace('.cu', '.cpp'))
        open(cpp, 'w').write(translate(open(os.path.join(CSRC, s)).read(), s))
        o = cpp.replace('.cpp', '.o')
        procs.append((s, subprocess.Popen([cxx] + flags + ['-c', cpp, '-o', o], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
        objs.append(o)
    rt = os.path.join(outdir, 'hostcc_rt.o')
    procs.append(('hostcc_rt.cpp', subprocess.Popen([cxx] + flags + ['-c', os.path.join(HERE, 'hostcc_rt.cpp'), '-o', rt], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    objs.append(rt)
    for s, p in procs:
        out, _ = p.communicate()
        if p.returncode != 0:
            raise RuntimeError(f'hostcc: {s} does not compile for the host:\n{out[-6000:]}')
    so = os.path.join(outdir, 'libchord_b200_hostcc_scratch.so')
    subprocess.check_call([cxx, '-shared', '-o', so] + objs + list(extra) + ['-ldl'])
    return so


if __name__ == '__main__':
    import sys
    print(build(sys.argv[1] if len(sys.argv) > 1 else '/tmp/hostcc_build', extra=sys.argv[2:]))
