"""Parser for the C subset that libnomp kernels are written in.

The reference parses kernel strings with libclang and maps a subset of C onto loopy
(reference python/loopy_api.py:156-292 expressions, :294-378 for-loops, :447-709 statements, :712-821 kernel).
Neither libclang nor loopy exist here, so this is a self-contained recursive-descent parser.  The accepted
language is a superset of the reference's: exactly one function (or a bare statement list, which gets wrapped),
scalar and pointer/array parameters, `for (T i = lo; i < hi; i++)` loops, if/else, declarations of scalars and
fixed/variable-length N-D arrays, `=` and compound assignments, break/continue, and C expressions with
arithmetic / comparison / logical / bitwise operators, subscripts, the ternary operator, casts, unary minus and
calls to math functions (the last four are extensions; the reference rejects them at python/loopy_api.py:45-48).

Pure Python, stdlib only: libnomp.so embeds CPython and must survive Py_FinalizeEx + re-initialisation in one
process (reference tests/nomp-api-020.c:3-27), which C-extension modules such as numpy do not.
"""
from __future__ import annotations

import re
from dataclasses import dataclass, field
from typing import List, Optional, Tuple

# ----------------------------------------------------------------------------------------------------
# AST
# ----------------------------------------------------------------------------------------------------


@dataclass
class CType:
    base: str          # bool char short int long longlong float double
    unsigned: bool = False
    const: bool = False
    ptr: int = 0       # pointer depth

    def scalar(self) -> "CType":
        return CType(self.base, self.unsigned, False, 0)

    def spell(self) -> str:
        b = {"longlong": "long long"}.get(self.base, self.base)
        s = ("unsigned " if self.unsigned else "") + b
        if self.const:
            s = "const " + s
        return s + " " + "*" * self.ptr if self.ptr else s

    @property
    def is_float(self) -> bool:
        return self.base in ("float", "double")

    @property
    def size(self) -> int:
        return {"bool": 1, "char": 1, "short": 2, "int": 4, "long": 8, "longlong": 8, "float": 4, "double": 8}[self.base]


@dataclass
class Node:
    pass


@dataclass
class Num(Node):
    text: str

    @property
    def is_int(self) -> bool:
        return re.fullmatch(r"(0[xX][0-9a-fA-F]+|\d+)[uUlL]*", self.text) is not None

    @property
    def value(self) -> int:
        return int(re.sub(r"[uUlL]+$", "", self.text), 0)


@dataclass
class Name(Node):
    id: str


@dataclass
class Subscript(Node):
    base: Node
    index: List[Node]


@dataclass
class BinOp(Node):
    op: str
    left: Node
    right: Node


@dataclass
class UnOp(Node):
    op: str
    operand: Node


@dataclass
class Ternary(Node):
    cond: Node
    then: Node
    other: Node


@dataclass
class Call(Node):
    func: str
    args: List[Node]


@dataclass
class Cast(Node):
    ctype: CType
    operand: Node


@dataclass
class Assign(Node):
    target: Node
    op: str            # "=", "+=", "-=", "*=", "/=", "%=", "&=", "|=", "^=", "<<=", ">>="
    value: Node


@dataclass
class Decl(Node):
    ctype: CType
    name: str
    dims: List[Node] = field(default_factory=list)
    init: Optional[Node] = None


@dataclass
class For(Node):
    var: str
    vtype: CType
    lo: Node
    hi: Node           # exclusive upper bound (a `<=` loop is stored with hi + 1)
    body: List[Node]
    tag: Optional[str] = None       # None | "for" | "g.N" | "l.N" (set by loopy.tag_inames)


@dataclass
class Bind(Node):
    """Result of split_iname: `var = lo + outer * size + inner`, body runs only when var < hi."""
    var: str
    vtype: CType
    value: Node
    hi: Node
    body: List[Node]


@dataclass
class If(Node):
    cond: Node
    then: List[Node]
    other: List[Node] = field(default_factory=list)


@dataclass
class Break(Node):
    pass


@dataclass
class Continue(Node):
    pass


@dataclass
class Param(Node):
    ctype: CType
    name: str
    dims: List[Node] = field(default_factory=list)

    @property
    def is_array(self) -> bool:
        return self.ctype.ptr > 0 or bool(self.dims)


@dataclass
class Function(Node):
    name: str
    params: List[Param]
    body: List[Node]


# ----------------------------------------------------------------------------------------------------
# lexer
# ----------------------------------------------------------------------------------------------------
_TOKEN_RE = re.compile(r"""
    (?P<ws>\s+|//[^\n]*|/\*.*?\*/)
  | (?P<num>(?:0[xX][0-9a-fA-F]+[uUlL]*)|(?:(?:\d+\.\d*|\.\d+|\d+)(?:[eE][+-]?\d+)?[fFuUlL]*))
  | (?P<id>[A-Za-z_][A-Za-z_0-9]*)
  | (?P<op><<=|>>=|\+\+|--|<<|>>|<=|>=|==|!=|&&|\|\||\+=|-=|\*=|/=|%=|&=|\|=|\^=|[-+*/%<>=!~&|^?:;,.(){}\[\]])
""", re.VERBOSE | re.DOTALL)

_TYPE_WORDS = {"void", "bool", "_Bool", "char", "short", "int", "long", "float", "double", "unsigned", "signed",
               "const", "volatile", "restrict", "__restrict__", "__restrict", "size_t", "static", "register"}

ASSIGN_OPS = {"=", "+=", "-=", "*=", "/=", "%=", "&=", "|=", "^=", "<<=", ">>="}

_BINARY_PRECEDENCE = [
    ("||",), ("&&",), ("|",), ("^",), ("&",), ("==", "!="), ("<", "<=", ">", ">="), ("<<", ">>"), ("+", "-"),
    ("*", "/", "%"),
]


class CSyntaxError(SyntaxError):
    pass


def tokenize(src: str) -> List[Tuple[str, str, int]]:
    toks, pos = [], 0
    while pos < len(src):
        m = _TOKEN_RE.match(src, pos)
        if not m:
            raise CSyntaxError(f"unexpected character {src[pos]!r} at offset {pos}")
        pos = m.end()
        kind = m.lastgroup
        if kind != "ws":
            toks.append((kind, m.group(kind), m.start()))
    toks.append(("eof", "", len(src)))
    return toks


class Parser:
    def __init__(self, src: str):
        self.src = src
        self.toks = tokenize(src)
        self.i = 0

    # -- token helpers -------------------------------------------------------------------------------
    def peek(self, k=0):
        return self.toks[min(self.i + k, len(self.toks) - 1)]

    def at(self, text: str, k=0) -> bool:
        t = self.peek(k)
        return t[1] == text and t[0] in ("op", "id")

    def next(self):
        t = self.toks[self.i]
        self.i += 1
        return t

    def expect(self, text: str):
        t = self.next()
        if t[1] != text:
            raise CSyntaxError(f"expected {text!r} but found {t[1]!r} at offset {t[2]}")
        return t

    def accept(self, text: str) -> bool:
        if self.at(text):
            self.i += 1
            return True
        return False

    def ident(self) -> str:
        t = self.next()
        if t[0] != "id" or t[1] in _TYPE_WORDS:
            raise CSyntaxError(f"expected an identifier but found {t[1]!r} at offset {t[2]}")
        return t[1]

    # -- types ---------------------------------------------------------------------------------------
    def at_type(self, k=0) -> bool:
        t = self.peek(k)
        return t[0] == "id" and t[1] in _TYPE_WORDS

    def parse_type(self) -> CType:
        words = []
        while self.at_type():
            words.append(self.next()[1])
        if not words:
            t = self.peek()
            raise CSyntaxError(f"expected a type but found {t[1]!r} at offset {t[2]}")
        const = "const" in words
        unsigned = "unsigned" in words
        core = [w for w in words if w not in ("const", "volatile", "restrict", "__restrict__", "__restrict", "unsigned",
                                               "signed", "static", "register")]
        if core == ["size_t"]:
            base, unsigned = "long", True
        elif core in (["bool"], ["_Bool"]):
            base = "bool"
        elif core == ["void"]:
            base = "void"
        elif core.count("long") == 2:
            base = "longlong"
        elif "long" in core and "double" not in core:
            base = "long"
        elif "short" in core:
            base = "short"
        elif "char" in core:
            base = "char"
        elif "float" in core:
            base = "float"
        elif "double" in core:
            base = "double"
        elif core in ([], ["int"]):
            base = "int"
        else:
            raise CSyntaxError(f"unsupported type {' '.join(words)!r}")
        ptr = 0
        while self.at("*"):
            self.next()
            ptr += 1
            while self.at_type() and self.peek()[1] in ("const", "restrict", "__restrict__", "__restrict", "volatile"):
                self.next()
        return CType(base, unsigned, const, ptr)

    # -- expressions ---------------------------------------------------------------------------------
    def parse_expr(self) -> Node:
        return self.parse_ternary()

    def parse_ternary(self) -> Node:
        cond = self.parse_binary(0)
        if self.accept("?"):
            a = self.parse_expr()
            self.expect(":")
            b = self.parse_ternary()
            return Ternary(cond, a, b)
        return cond

    def parse_binary(self, level: int) -> Node:
        if level == len(_BINARY_PRECEDENCE):
            return self.parse_unary()
        left = self.parse_binary(level + 1)
        ops = _BINARY_PRECEDENCE[level]
        while self.peek()[0] == "op" and self.peek()[1] in ops:
            op = self.next()[1]
            right = self.parse_binary(level + 1)
            left = BinOp(op, left, right)
        return left

    def parse_unary(self) -> Node:
        t = self.peek()
        if t[0] == "op" and t[1] in ("-", "+", "!", "~"):
            self.next()
            return UnOp(t[1], self.parse_unary())
        if t[0] == "op" and t[1] in ("++", "--"):
            raise CSyntaxError(f"{t[1]} is only supported as the update of a for loop (offset {t[2]})")
        if t[1] == "(" and self.at_type(1):
            self.next()
            ct = self.parse_type()
            self.expect(")")
            return Cast(ct, self.parse_unary())
        return self.parse_postfix()

    def parse_postfix(self) -> Node:
        t = self.next()
        if t[0] == "num":
            node: Node = Num(t[1])
        elif t[0] == "id" and t[1] not in _TYPE_WORDS:
            if self.at("("):
                self.next()
                args = []
                if not self.at(")"):
                    args.append(self.parse_expr())
                    while self.accept(","):
                        args.append(self.parse_expr())
                self.expect(")")
                node = Call(t[1], args)
            else:
                node = Name(t[1])
        elif t[1] == "(":
            node = self.parse_expr()
            self.expect(")")
        else:
            raise CSyntaxError(f"unexpected token {t[1]!r} at offset {t[2]}")
        idx = []
        while self.at("["):
            self.next()
            idx.append(self.parse_expr())
            self.expect("]")
        if idx:
            node = Subscript(node, idx)
        if self.peek()[0] == "op" and self.peek()[1] in ("++", "--"):
            t = self.peek()
            raise CSyntaxError(f"{t[1]} is only supported as the update of a for loop (offset {t[2]})")
        return node

    # -- statements ----------------------------------------------------------------------------------
    def parse_block_or_stmt(self) -> List[Node]:
        if self.at("{"):
            return self.parse_compound()
        return [self.parse_stmt()]

    def parse_compound(self) -> List[Node]:
        self.expect("{")
        out = []
        while not self.at("}"):
            if self.peek()[0] == "eof":
                raise CSyntaxError("unexpected end of kernel source: missing '}'")
            out.append(self.parse_stmt())
        self.expect("}")
        return out

    def parse_stmt(self) -> Node:
        t = self.peek()
        if t[1] == "{":
            body = self.parse_compound()
            return If(Num("1"), body) if len(body) != 1 else body[0]
        if t[1] == ";":
            self.next()
            return If(Num("1"), [])
        if t[0] == "id" and t[1] == "for":
            return self.parse_for()
        if t[0] == "id" and t[1] == "if":
            self.next()
            self.expect("(")
            cond = self.parse_expr()
            self.expect(")")
            then = self.parse_block_or_stmt()
            other = []
            if self.peek()[0] == "id" and self.peek()[1] == "else":
                self.next()
                other = self.parse_block_or_stmt()
            return If(cond, then, other)
        if t[0] == "id" and t[1] == "break":
            self.next()
            self.expect(";")
            return Break()
        if t[0] == "id" and t[1] == "continue":
            self.next()
            self.expect(";")
            return Continue()
        if t[0] == "id" and t[1] in ("while", "do", "switch", "goto", "return"):
            raise CSyntaxError(f"'{t[1]}' statements are not supported in nomp kernels (offset {t[2]})")
        if self.at_type():
            return self.parse_decl()
        target = self.parse_postfix()
        op = self.next()
        if op[1] not in ASSIGN_OPS:
            raise CSyntaxError(f"expected an assignment operator but found {op[1]!r} at offset {op[2]}")
        value = self.parse_expr()
        self.expect(";")
        if not isinstance(target, (Name, Subscript)):
            raise CSyntaxError("left-hand side of an assignment must be a variable or an array element")
        return Assign(target, op[1], value)

    def parse_decl(self) -> Node:
        ct = self.parse_type()
        name = self.ident()
        dims = []
        while self.accept("["):
            dims.append(self.parse_expr())
            self.expect("]")
        init = None
        if self.accept("="):
            init = self.parse_expr()
        if self.at(","):
            raise CSyntaxError("one variable per declaration, please")
        self.expect(";")
        return Decl(ct, name, dims, init)

    def parse_for(self) -> Node:
        self.expect("for")
        self.expect("(")
        if not self.at_type():
            raise CSyntaxError("the loop variable must be declared in the for-loop initialiser")
        vt = self.parse_type()
        var = self.ident()
        self.expect("=")
        lo = self.parse_expr()
        if self.at(","):
            raise CSyntaxError("multiple loop variables are not supported")
        self.expect(";")
        cv = self.ident()
        cmp_ = self.next()
        if cmp_[1] not in ("<", "<="):
            raise CSyntaxError("the for-loop condition must use < or <=")
        hi = self.parse_expr()
        if cmp_[1] == "<=":
            hi = BinOp("+", hi, Num("1"))
        self.expect(";")
        # i++ | ++i | i += 1
        if self.at("++"):
            self.next()
            uv = self.ident()
        else:
            uv = self.ident()
            if self.at("++"):
                self.next()
            elif self.at("+=") and self.peek(1)[1] == "1":
                self.next()
                self.next()
            else:
                raise CSyntaxError("the for-loop update must be ++")
        if cv != var or uv != var:
            raise CSyntaxError("the loop variable must be the same in initialiser, condition and update")
        self.expect(")")
        body = self.parse_block_or_stmt()
        return For(var, vt, lo, hi, body)

    # -- top level -----------------------------------------------------------------------------------
    def parse_function(self) -> Function:
        # `void name(params) { ... }`
        if not self.at_type():
            raise CSyntaxError("a kernel must be a single C function")
        self.parse_type()
        name = self.ident()
        self.expect("(")
        params = []
        if not self.at(")"):
            while True:
                if self.at("void") and self.peek(1)[1] == ")":
                    self.next()
                    break
                ct = self.parse_type()
                pn = self.ident()
                dims = []
                while self.accept("["):
                    if self.at("]"):
                        dims.append(None)
                    else:
                        dims.append(self.parse_expr())
                    self.expect("]")
                if dims:
                    ct = CType(ct.base, ct.unsigned, ct.const, ct.ptr + 1)
                params.append(Param(ct, pn, dims))
                if not self.accept(","):
                    break
        self.expect(")")
        body = self.parse_compound()
        if self.peek()[0] != "eof":
            t = self.peek()
            raise CSyntaxError(f"unexpected text after the kernel function at offset {t[2]}")
        return Function(name, params, body)


def parse_kernel(src: str) -> Function:
    """Parse a kernel string into a Function.  Raises CSyntaxError."""
    return Parser(src).parse_function()
