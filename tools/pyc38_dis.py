#!/usr/bin/env python3
"""Stand-alone disassembler for CPython 3.8 ``.pyc`` files (runs on any Python 3).

Why it exists: two of the three CLR losses (the prototype-guided discriminative
hinge and the augmented-consistency BCE) ship in the reference ONLY as bytecode
(``train_process/__pycache__/Trainer_prototype_mt.cpython-38.pyc``).  Python
3.12's ``marshal``/``dis`` cannot read 3.8 code objects, so this file carries
its own marshal reader and the 3.8 opcode table.  It is an investigation tool:
nothing in the product, the tests or the bench imports it.

Usage:
    python tools/pyc38_dis.py FILE.pyc [--func NAME] [--lines LO:HI]
"""
import struct
import sys

OPNAMES = {
    1: "POP_TOP", 2: "ROT_TWO", 3: "ROT_THREE", 4: "DUP_TOP", 5: "DUP_TOP_TWO", 6: "ROT_FOUR", 9: "NOP",
    10: "UNARY_POSITIVE", 11: "UNARY_NEGATIVE", 12: "UNARY_NOT", 15: "UNARY_INVERT",
    16: "BINARY_MATRIX_MULTIPLY", 17: "INPLACE_MATRIX_MULTIPLY", 19: "BINARY_POWER", 20: "BINARY_MULTIPLY",
    22: "BINARY_MODULO", 23: "BINARY_ADD", 24: "BINARY_SUBTRACT", 25: "BINARY_SUBSCR",
    26: "BINARY_FLOOR_DIVIDE", 27: "BINARY_TRUE_DIVIDE", 28: "INPLACE_FLOOR_DIVIDE", 29: "INPLACE_TRUE_DIVIDE",
    50: "GET_AITER", 51: "GET_ANEXT", 52: "BEFORE_ASYNC_WITH", 53: "BEGIN_FINALLY", 54: "END_ASYNC_FOR",
    55: "INPLACE_ADD", 56: "INPLACE_SUBTRACT", 57: "INPLACE_MULTIPLY", 59: "INPLACE_MODULO",
    60: "STORE_SUBSCR", 61: "DELETE_SUBSCR", 62: "BINARY_LSHIFT", 63: "BINARY_RSHIFT", 64: "BINARY_AND",
    65: "BINARY_XOR", 66: "BINARY_OR", 67: "INPLACE_POWER", 68: "GET_ITER", 69: "GET_YIELD_FROM_ITER",
    70: "PRINT_EXPR", 71: "LOAD_BUILD_CLASS", 72: "YIELD_FROM", 73: "GET_AWAITABLE", 75: "INPLACE_LSHIFT",
    76: "INPLACE_RSHIFT", 77: "INPLACE_AND", 78: "INPLACE_XOR", 79: "INPLACE_OR", 81: "WITH_CLEANUP_START",
    82: "WITH_CLEANUP_FINISH", 83: "RETURN_VALUE", 84: "IMPORT_STAR", 85: "SETUP_ANNOTATIONS",
    86: "YIELD_VALUE", 87: "POP_BLOCK", 88: "END_FINALLY", 89: "POP_EXCEPT",
    90: "STORE_NAME", 91: "DELETE_NAME", 92: "UNPACK_SEQUENCE", 93: "FOR_ITER", 94: "UNPACK_EX",
    95: "STORE_ATTR", 96: "DELETE_ATTR", 97: "STORE_GLOBAL", 98: "DELETE_GLOBAL", 100: "LOAD_CONST",
    101: "LOAD_NAME", 102: "BUILD_TUPLE", 103: "BUILD_LIST", 104: "BUILD_SET", 105: "BUILD_MAP",
    106: "LOAD_ATTR", 107: "COMPARE_OP", 108: "IMPORT_NAME", 109: "IMPORT_FROM", 110: "JUMP_FORWARD",
    111: "JUMP_IF_FALSE_OR_POP", 112: "JUMP_IF_TRUE_OR_POP", 113: "JUMP_ABSOLUTE", 114: "POP_JUMP_IF_FALSE",
    115: "POP_JUMP_IF_TRUE", 116: "LOAD_GLOBAL", 122: "SETUP_FINALLY", 124: "LOAD_FAST", 125: "STORE_FAST",
    126: "DELETE_FAST", 130: "RAISE_VARARGS", 131: "CALL_FUNCTION", 132: "MAKE_FUNCTION", 133: "BUILD_SLICE",
    135: "LOAD_CLOSURE", 136: "LOAD_DEREF", 137: "STORE_DEREF", 138: "DELETE_DEREF", 141: "CALL_FUNCTION_KW",
    142: "CALL_FUNCTION_EX", 143: "SETUP_WITH", 144: "EXTENDED_ARG", 145: "LIST_APPEND", 146: "SET_ADD",
    147: "MAP_ADD", 148: "LOAD_CLASSDEREF", 149: "BUILD_LIST_UNPACK", 150: "BUILD_MAP_UNPACK",
    151: "BUILD_MAP_UNPACK_WITH_CALL", 152: "BUILD_TUPLE_UNPACK", 153: "BUILD_SET_UNPACK",
    154: "SETUP_ASYNC_WITH", 155: "FORMAT_VALUE", 156: "BUILD_CONST_KEY_MAP", 157: "BUILD_STRING",
    158: "BUILD_TUPLE_UNPACK_WITH_CALL", 160: "LOAD_METHOD", 161: "CALL_METHOD", 162: "CALL_FINALLY",
    163: "POP_FINALLY",
}
CMP = ["<", "<=", "==", "!=", ">", ">=", "in", "not in", "is", "is not", "exception match", "BAD"]
HAS_CONST = {100}
HAS_NAME = {90, 91, 95, 96, 97, 98, 101, 106, 108, 109, 116, 160}
HAS_LOCAL = {124, 125, 126}
HAS_FREE = {135, 136, 137, 138, 148}


class Code:
    pass


class Reader:
    def __init__(self, data):
        self.d = data
        self.p = 0
        self.refs = []

    def u8(self):
        v = self.d[self.p]
        self.p += 1
        return v

    def i32(self):
        v = struct.unpack_from("<i", self.d, self.p)[0]
        self.p += 4
        return v

    def raw(self, n):
        v = self.d[self.p:self.p + n]
        self.p += n
        return v

    def obj(self):
        t = self.u8()
        flag = t & 0x80
        t = chr(t & 0x7F)
        idx = None
        if flag:
            idx = len(self.refs)
            self.refs.append(None)

        def keep(v):
            if idx is not None:
                self.refs[idx] = v
            return v

        if t == "0":
            return None
        if t == "N":
            return None
        if t == "T":
            return True
        if t == "F":
            return False
        if t == ".":
            return Ellipsis
        if t == "S":
            return StopIteration
        if t == "i":
            return keep(self.i32())
        if t == "l":
            n = self.i32()
            digs = [struct.unpack_from("<H", self.raw(2))[0] for _ in range(abs(n))]
            v = sum(dg << (15 * k) for k, dg in enumerate(digs))
            return keep(-v if n < 0 else v)
        if t == "g":
            return keep(struct.unpack("<d", self.raw(8))[0])
        if t == "y":
            re, im = struct.unpack("<dd", self.raw(16))
            return keep(complex(re, im))
        if t == "s":
            return keep(bytes(self.raw(self.i32())))
        if t in "ut":
            return keep(self.raw(self.i32()).decode("utf-8", "surrogatepass"))
        if t in "aA":
            return keep(self.raw(self.i32()).decode("latin-1"))
        if t in "zZ":
            return keep(self.raw(self.u8()).decode("latin-1"))
        if t == ")":
            n = self.u8()
            return keep(tuple(self.obj() for _ in range(n)))
        if t == "(":
            n = self.i32()
            return keep(tuple(self.obj() for _ in range(n)))
        if t == "[":
            n = self.i32()
            return keep([self.obj() for _ in range(n)])
        if t in "<>":
            n = self.i32()
            v = [self.obj() for _ in range(n)]
            return keep(frozenset(v))
        if t == "{":
            out = {}
            while True:
                k = self.obj()
                if k is None and self.d[self.p - 1] == ord("0"):
                    break
                out[k] = self.obj()
            return keep(out)
        if t == "r":
            return self.refs[self.i32()]
        if t == "c":
            c = Code()
            keep(c)
            c.argcount = self.i32()
            c.posonly = self.i32()
            c.kwonly = self.i32()
            c.nlocals = self.i32()
            c.stacksize = self.i32()
            c.flags = self.i32()
            c.code = self.obj()
            c.consts = self.obj()
            c.names = self.obj()
            c.varnames = self.obj()
            c.freevars = self.obj()
            c.cellvars = self.obj()
            c.filename = self.obj()
            c.name = self.obj()
            c.firstlineno = self.i32()
            c.lnotab = self.obj()
            return c
        raise ValueError("unknown marshal type %r at %d" % (t, self.p - 1))


def line_table(c):
    out = {}
    line = c.firstlineno
    addr = 0
    out[0] = line
    tab = c.lnotab
    for i in range(0, len(tab), 2):
        da, dl = tab[i], tab[i + 1]
        if dl >= 128:
            dl -= 256
        addr += da
        line += dl
        out[addr] = line
    return out


def show(v):
    if isinstance(v, Code):
        return "<code %s>" % v.name
    return repr(v)


def dis(c, lo=None, hi=None, out=sys.stdout):
    lt = line_table(c)
    code = c.code
    ext = 0
    cur = None
    for off in range(0, len(code), 2):
        op, arg = code[off], code[off + 1]
        if off in lt:
            cur = lt[off]
        if op == 144:
            ext = (ext | arg) << 8
            continue
        arg |= ext
        ext = 0
        if lo is not None and (cur is None or cur < lo or cur > hi):
            continue
        name = OPNAMES.get(op, "OP_%d" % op)
        extra = ""
        if op >= 90:
            if op in HAS_CONST:
                extra = show(c.consts[arg])
            elif op in HAS_NAME:
                extra = c.names[arg]
            elif op in HAS_LOCAL:
                extra = c.varnames[arg]
            elif op in HAS_FREE:
                extra = (c.cellvars + c.freevars)[arg]
            elif op == 107:
                extra = CMP[arg]
            else:
                extra = str(arg)
        out.write("L%-5s %5d %-28s %s\n" % (cur, off, name, extra))


def walk(c, fn):
    fn(c)
    for k in c.consts:
        if isinstance(k, Code):
            walk(k, fn)


def main(argv):
    path = argv[1]
    func = None
    lo = hi = None
    i = 2
    while i < len(argv):
        if argv[i] == "--func":
            func = argv[i + 1]
            i += 2
        elif argv[i] == "--lines":
            lo, hi = (int(x) for x in argv[i + 1].split(":"))
            i += 2
        else:
            raise SystemExit("unknown arg " + argv[i])
    data = open(path, "rb").read()
    top = Reader(data[16:]).obj()

    def visit(c):
        if func is None or c.name == func:
            print("== %s (line %d) args=%s" % (c.name, c.firstlineno, c.varnames[:c.argcount]))
            if func is not None or lo is not None:
                dis(c, lo, hi)

    walk(top, visit)


if __name__ == "__main__":
    main(sys.argv)
