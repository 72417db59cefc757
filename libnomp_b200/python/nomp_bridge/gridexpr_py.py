"""Python mirror of src/gridexpr.c (launch-size expression evaluator); used by tests and tooling."""
from __future__ import annotations

import re

_TOK = re.compile(r"\s*(\d+|[A-Za-z_]\w*|[-+*/%(),])")


def evaluate(expr: str, env: dict) -> int:
    toks = []
    pos = 0
    while pos < len(expr):
        if expr[pos:].strip() == "":
            break
        m = _TOK.match(expr, pos)
        if not m:
            raise ValueError(f"bad launch-size expression: {expr!r}")
        toks.append(m.group(1))
        pos = m.end()
    i = 0

    def peek():
        return toks[i] if i < len(toks) else None

    def take():
        nonlocal i
        i += 1
        return toks[i - 1]

    def factor():
        t = take()
        if t == "(":
            v = expression()
            if take() != ")":
                raise ValueError("expected )")
            return v
        if t == "-":
            return -factor()
        if t == "+":
            return factor()
        if t.isdigit():
            return int(t)
        if t in ("min", "max") and peek() == "(":
            take()
            a = expression()
            if take() != ",":
                raise ValueError("expected ,")
            b = expression()
            if take() != ")":
                raise ValueError("expected )")
            return min(a, b) if t == "min" else max(a, b)
        if t in env:
            return int(env[t])
        raise ValueError(f"unknown name {t!r}")

    def term():
        v = factor()
        while peek() in ("*", "/", "%"):
            op = take()
            r = factor()
            if op == "*":
                v *= r
            elif r == 0:
                raise ZeroDivisionError
            elif op == "/":
                v = int(v / r)        # truncation, like C
            else:
                v = v - r * int(v / r)
        return v

    def expression():
        v = term()
        while peek() in ("+", "-"):
            op = take()
            r = term()
            v = v + r if op == "+" else v - r
        return v

    out = expression()
    if i != len(toks):
        raise ValueError("trailing tokens")
    return out
