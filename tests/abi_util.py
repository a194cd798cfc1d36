"""Parse include/esb200.h into ctypes-style argument codes (test helper)."""
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _code(decl):
    decl = decl.strip()
    if decl in ('void', ''):
        return ''
    if '*' in decl:
        return 'p'
    base = re.sub(r'\b(const|unsigned)\b', '', decl).split()
    ty = ' '.join(base[:-1]) if len(base) > 1 else base[0]
    ty = ty.strip()
    return {'long long': 'q', 'int': 'i', 'float': 'f', 'size_t': 'z'}[ty]


def header_signatures():
    text = open(os.path.join(ROOT, 'include', 'esb200.h')).read()
    text = re.sub(r'/\*.*?\*/', '', text, flags=re.S)
    sigs = {}
    for m in re.finditer(r'([\w\s\*]+?)\b(esb_\w+)\s*\(([^;{]*?)\)\s*;', text):
        ret, name, args = m.group(1).strip(), m.group(2), m.group(3)
        rc = 'p' if '*' in ret else {'int': 'i', 'long long': 'q', 'size_t': 'z'}[ret]
        sigs[name] = (''.join(_code(a) for a in args.split(',')), rc)
    return sigs
