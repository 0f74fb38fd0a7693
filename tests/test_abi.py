"""The C-ABI: header <-> ctypes table <-> exported symbols (CPU only, no compute calls)."""
import ctypes
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, 'include', 'fsdet.h')
LIB = os.path.join(ROOT, 'fewshot_detection_b200', 'libfsdet.so')


def parse_header():
    src = open(HEADER).read()
    src = re.sub(r'/\*.*?\*/', ' ', src, flags=re.S)
    decls = {}
    for m in re.finditer(r'(?:^|\n)\s*(int|size_t|const char\*)\s+(fsdet_\w+)\s*\(([^;]*?)\)\s*;', src):
        ret, name, params = m.group(1), m.group(2), m.group(3).strip()
        codes = ''
        if params and params != 'void':
            for prm in params.split(','):
                prm = prm.strip()
                if '*' in prm:
                    codes += 'p'
                elif 'size_t' in prm:
                    codes += 'z'
                elif 'long long' in prm:
                    codes += 'q'
                elif 'double' in prm:
                    codes += 'd'
                elif 'float' in prm:
                    codes += 'f'
                elif 'int' in prm:
                    codes += 'i'
                else:
                    raise AssertionError('unparsed parameter %r in %s' % (prm, name))
        decls[name] = (codes, {'int': 'i', 'size_t': 'z', 'const char*': 's'}[ret])
    return decls


def _build_if_needed():
    if not os.path.exists(LIB):
        import sys
        sys.path.insert(0, ROOT)
        import __graft_entry__
        __graft_entry__.build()


def test_header_matches_ctypes_table():
    _build_if_needed()
    from fewshot_detection_b200 import _lib
    decls = parse_header()
    assert len(decls) >= 30
    assert set(decls) == set(_lib.SIGNATURES), set(decls) ^ set(_lib.SIGNATURES)
    for name, sig in decls.items():
        assert _lib.SIGNATURES[name] == sig, (name, _lib.SIGNATURES[name], sig)


def test_library_exports_every_declared_symbol():
    _build_if_needed()
    lib = ctypes.CDLL(LIB)
    for name in parse_header():
        assert hasattr(lib, name), name
    lib.fsdet_version.restype = ctypes.c_int
    assert lib.fsdet_version() >= 100
    assert lib.fsdet_compiled_arch() == 100


def test_library_is_sm100a_native():
    _build_if_needed()
    out = subprocess.run(['cuobjdump', '-lelf', LIB], capture_output=True, text=True).stdout
    assert 'sm_100a' in out, out


def test_product_package_never_imports_the_oracle():
    """oracle/ is test infrastructure: only tests/, __graft_entry__.smoke() and bench.py's CPU legs may import it."""
    pkg = os.path.join(ROOT, 'fewshot_detection_b200')
    offenders = []
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith(('.py', '.cu', '.cuh', '.h')):
                src = open(os.path.join(dirpath, f)).read()
                if re.search(r'^\s*(from|import)\s+oracle\b', src, flags=re.M):
                    offenders.append(os.path.join(dirpath, f))
    assert offenders == []
    for f in os.listdir(os.path.join(ROOT, 'dropin')):
        assert 'oracle' not in open(os.path.join(ROOT, 'dropin', f)).read()
