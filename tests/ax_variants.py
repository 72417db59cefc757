"""Spellings of the spectral-element Ax operator other than the canonical kernel string of nomp_bridge/families.py, for
the structural recogniser (nomp_bridge/axprobe.py).  Each is plain nomp C with `n` passed as NOMP_INT | NOMP_JIT; the
tests compile them with gcc against the oracle (so that the texts themselves are known to be the operator) and require the
bridge to route every one of them to the hand-written kernel.  NOT_AX are near misses that must keep the generic path.

(name, source, argument order as roles) -- roles: w u g D E n [pap]
"""

POINT = "e * n * n * n + k * n * n + j * n + i"

# 1. other names, g as a five-dimensional and D as a two-dimensional array, flat temporaries, statements in another
#    order, integer temporaries for the subscripts, w zeroed and then accumulated into with +=
V_MULTIDIM = ("multidim", """
void my_ax(int E, int n, double *out, const double *in, const double g[E][6][n][n][n], const double dm[n][n]) {
  for (int e = 0; e < E; e++) {
    double wr[n * n * n];
    double ws[n * n * n];
    double wt[n * n * n];
    for (int k = 0; k < n; k++)
      for (int j = 0; j < n; j++)
        for (int i = 0; i < n; i++) {
          int p = k * n * n + j * n + i;
          int base = e * n * n * n;
          double t = 0;
          double s = 0;
          double r = 0;
          for (int l = 0; l < n; l++) {
            t += dm[k][l] * in[base + l * n * n + j * n + i];
            r += dm[i][l] * in[base + k * n * n + j * n + l];
            s += dm[j][l] * in[base + k * n * n + l * n + i];
          }
          wt[p] = g[e][2][k][j][i] * r + g[e][4][k][j][i] * s + g[e][5][k][j][i] * t;
          wr[p] = g[e][0][k][j][i] * r + g[e][1][k][j][i] * s + g[e][2][k][j][i] * t;
          ws[p] = g[e][1][k][j][i] * r + g[e][3][k][j][i] * s + g[e][4][k][j][i] * t;
          out[base + p] = 0;
        }
    for (int k = 0; k < n; k++)
      for (int j = 0; j < n; j++)
        for (int i = 0; i < n; i++)
          for (int l = 0; l < n; l++) {
            out[e * n * n * n + k * n * n + j * n + i] += dm[l][k] * wt[l * n * n + j * n + i];
            out[e * n * n * n + k * n * n + j * n + i] += dm[l][i] * wr[k * n * n + j * n + l] + dm[l][j] * ws[k * n * n + l * n + i];
          }
  }
}
""", ("out", "in", "g", "dm", "E", "n"))

# 2. the point loops in the opposite order (i outermost), arguments in another order, the three transposed contractions in
#    three separate loops over l
V_LOOP_ORDER = ("loop_order", f"""
void ax_ikj(const double *D, const double *g, const double *u, double *w, int n, int E) {{
  for (int e = 0; e < E; e++) {{
    double a[n][n][n];
    double b[n][n][n];
    double cc[n][n][n];
    for (int i = 0; i < n; i++)
      for (int k = 0; k < n; k++)
        for (int j = 0; j < n; j++) {{
          double ur = 0;
          for (int l = 0; l < n; l++) ur += D[i * n + l] * u[e * n * n * n + k * n * n + j * n + l];
          double us = 0;
          for (int l = 0; l < n; l++) us += D[j * n + l] * u[e * n * n * n + k * n * n + l * n + i];
          double ut = 0;
          for (int l = 0; l < n; l++) ut += D[k * n + l] * u[e * n * n * n + l * n * n + j * n + i];
          a[k][j][i] = g[(e * 6 + 0) * n * n * n + k * n * n + j * n + i] * ur + g[(e * 6 + 1) * n * n * n + k * n * n + j * n + i] * us + g[(e * 6 + 2) * n * n * n + k * n * n + j * n + i] * ut;
          b[k][j][i] = g[(e * 6 + 1) * n * n * n + k * n * n + j * n + i] * ur + g[(e * 6 + 3) * n * n * n + k * n * n + j * n + i] * us + g[(e * 6 + 4) * n * n * n + k * n * n + j * n + i] * ut;
          cc[k][j][i] = g[(e * 6 + 2) * n * n * n + k * n * n + j * n + i] * ur + g[(e * 6 + 4) * n * n * n + k * n * n + j * n + i] * us + g[(e * 6 + 5) * n * n * n + k * n * n + j * n + i] * ut;
        }}
    for (int i = 0; i < n; i++)
      for (int j = 0; j < n; j++)
        for (int k = 0; k < n; k++) {{
          double s1 = 0;
          double s2 = 0;
          double s3 = 0;
          for (int l = 0; l < n; l++) s3 += D[l * n + k] * cc[l][j][i];
          for (int l = 0; l < n; l++) s1 += D[l * n + i] * a[k][j][l];
          for (int l = 0; l < n; l++) s2 += D[l * n + j] * b[k][l][i];
          w[{POINT}] = s1 + s2 + s3;
        }}
  }}
}}
""", ("w", "u", "g", "D", "E", "n"))

# 3. one temporary array for the three fluxes, the gradient in a small array, a sign written twice
V_ONE_TEMP = ("one_temporary", f"""
void ax_flux(double *w, const double *u, const double *g, const double *D, int E, int n) {{
  for (int e = 0; e < E; e++) {{
    double flux[3][n][n][n];
    for (int k = 0; k < n; k++)
      for (int j = 0; j < n; j++)
        for (int i = 0; i < n; i++) {{
          double grad[3];
          grad[0] = 0;
          grad[1] = 0;
          grad[2] = 0;
          for (int l = 0; l < n; l++) {{
            grad[0] += D[i * n + l] * u[e * n * n * n + k * n * n + j * n + l];
            grad[1] -= -D[j * n + l] * u[e * n * n * n + k * n * n + l * n + i];
            grad[2] += u[e * n * n * n + l * n * n + j * n + i] * D[k * n + l];
          }}
          int gp = e * 6 * n * n * n + k * n * n + j * n + i;
          flux[0][k][j][i] = g[gp] * grad[0] + g[gp + n * n * n] * grad[1] + g[gp + 2 * n * n * n] * grad[2];
          flux[1][k][j][i] = g[gp + n * n * n] * grad[0] + g[gp + 3 * n * n * n] * grad[1] + g[gp + 4 * n * n * n] * grad[2];
          flux[2][k][j][i] = g[gp + 2 * n * n * n] * grad[0] + g[gp + 4 * n * n * n] * grad[1] + g[gp + 5 * n * n * n] * grad[2];
        }}
    for (int k = 0; k < n; k++)
      for (int j = 0; j < n; j++)
        for (int i = 0; i < n; i++) {{
          double acc = 0;
          for (int l = 0; l < n; l++)
            acc += D[l * n + i] * flux[0][k][j][l] + D[l * n + j] * flux[1][k][l][i] + D[l * n + k] * flux[2][l][j][i];
          w[{POINT}] = acc;
        }}
  }}
}}
""", ("w", "u", "g", "D", "E", "n"))

# 4. the gradient stored first (three passes over the element instead of two), w = ... written as a product with 1.0
V_THREE_PASSES = ("three_passes", f"""
void ax3(double *w, const double *u, const double *g, const double *D, int E, int n) {{
  for (int e = 0; e < E; e++) {{
    double ur[n][n][n];
    double us[n][n][n];
    double ut[n][n][n];
    double wr[n][n][n];
    double ws[n][n][n];
    double wt[n][n][n];
    for (int k = 0; k < n; k++)
      for (int j = 0; j < n; j++)
        for (int i = 0; i < n; i++) {{
          ur[k][j][i] = 0;
          us[k][j][i] = 0;
          ut[k][j][i] = 0;
          for (int l = 0; l < n; l++) {{
            ur[k][j][i] += D[i * n + l] * u[e * n * n * n + k * n * n + j * n + l];
            us[k][j][i] += D[j * n + l] * u[e * n * n * n + k * n * n + l * n + i];
            ut[k][j][i] += D[k * n + l] * u[e * n * n * n + l * n * n + j * n + i];
          }}
        }}
    for (int k = 0; k < n; k++)
      for (int j = 0; j < n; j++)
        for (int i = 0; i < n; i++) {{
          wr[k][j][i] = g[(e * 6 + 0) * n * n * n + k * n * n + j * n + i] * ur[k][j][i] + g[(e * 6 + 1) * n * n * n + k * n * n + j * n + i] * us[k][j][i] + g[(e * 6 + 2) * n * n * n + k * n * n + j * n + i] * ut[k][j][i];
          ws[k][j][i] = g[(e * 6 + 1) * n * n * n + k * n * n + j * n + i] * ur[k][j][i] + g[(e * 6 + 3) * n * n * n + k * n * n + j * n + i] * us[k][j][i] + g[(e * 6 + 4) * n * n * n + k * n * n + j * n + i] * ut[k][j][i];
          wt[k][j][i] = g[(e * 6 + 2) * n * n * n + k * n * n + j * n + i] * ur[k][j][i] + g[(e * 6 + 4) * n * n * n + k * n * n + j * n + i] * us[k][j][i] + g[(e * 6 + 5) * n * n * n + k * n * n + j * n + i] * ut[k][j][i];
        }}
    for (int k = 0; k < n; k++)
      for (int j = 0; j < n; j++)
        for (int i = 0; i < n; i++) {{
          double acc = 0.0;
          for (int l = 0; l < n; l++) {{
            acc += D[l * n + k] * wt[l][j][i];
            acc += D[l * n + j] * ws[k][l][i];
            acc += D[l * n + i] * wr[k][j][l];
          }}
          w[{POINT}] = 1.0 * acc;
        }}
  }}
}}
""", ("w", "u", "g", "D", "E", "n"))


def _canonical():
    import sys
    from pathlib import Path
    sys.path.insert(0, str(Path(__file__).resolve().parent.parent / "libnomp_b200" / "python"))
    from nomp_bridge.families import AX_DOT_KERNEL_SOURCE, AX_KERNEL_SOURCE
    return AX_KERNEL_SOURCE, AX_DOT_KERNEL_SOURCE


def variants():
    ax, ax_dot = _canonical()
    # 5. the canonical string with the three accumulations of the second phase in another order
    reordered = ax.replace("            acc += D[l * n + i] * ur[k][j][l];\n            acc += D[l * n + j] * us[k][l][i];\n",
                           "            acc += D[l * n + j] * us[k][l][i];\n            acc += D[l * n + i] * ur[k][j][l];\n")
    assert reordered != ax
    # 6. ... and with the geometric factors multiplied from the right
    commuted = ax.replace("g[(e * 6 + 0) * n * n * n + k * n * n + j * n + i] * r", "r * g[(e * 6 + 0) * n * n * n + k * n * n + j * n + i]")
    assert commuted != ax
    return [V_MULTIDIM, V_LOOP_ORDER, V_ONE_TEMP, V_THREE_PASSES,
            ("statement_order", reordered, ("w", "u", "g", "D", "E", "n")),
            ("commuted_product", commuted, ("w", "u", "g", "D", "E", "n"))]


def fused_variants():
    """Ax + p.Ap under a reduce clause on `pap`, spelled differently from AX_DOT_KERNEL_SOURCE."""
    _, ax_dot = _canonical()
    swapped = ax_dot.replace("          w[e * n * n * n + k * n * n + j * n + i] = acc;\n          pap[0] += u[e * n * n * n + k * n * n + j * n + i] * acc;\n",
                             "          pap[0] += acc * u[e * n * n * n + k * n * n + j * n + i];\n          w[e * n * n * n + k * n * n + j * n + i] = acc;\n")
    assert swapped != ax_dot
    three = V_THREE_PASSES[1].replace("void ax3(double *w, const double *u, const double *g, const double *D, int E, int n) {",
                                      "void ax3_dot(double *w, const double *u, const double *g, const double *D, int E, int n, double *pap) {").replace(
        f"          w[{POINT}] = 1.0 * acc;\n", f"          w[{POINT}] = 1.0 * acc;\n          pap[0] += acc * u[{POINT}];\n")
    assert "pap[0]" in three
    return [("dot_statement_order", swapped, ("w", "u", "g", "D", "E", "n", "pap")),
            ("dot_three_passes", three, ("w", "u", "g", "D", "E", "n", "pap"))]


def not_ax():
    """Near misses: each differs from the operator in one place and must NOT reach the native kernel."""
    ax, _ = _canonical()
    out = [("wrong_transpose", ax.replace("acc += D[l * n + k] * ut[l][j][i];", "acc += D[k * n + l] * ut[l][j][i];")),
           ("factors_swapped", ax.replace("g[(e * 6 + 1) * n * n * n + k * n * n + j * n + i] * s + g[(e * 6 + 2)",
                                          "g[(e * 6 + 2) * n * n * n + k * n * n + j * n + i] * s + g[(e * 6 + 1)")),
           ("accumulates_into_w", ax.replace("w[e * n * n * n + k * n * n + j * n + i] = acc;", "w[e * n * n * n + k * n * n + j * n + i] += acc;")),
           ("scaled", ax.replace("w[e * n * n * n + k * n * n + j * n + i] = acc;", "w[e * n * n * n + k * n * n + j * n + i] = 2 * acc;")),
           ("branch", ax.replace("double acc = 0;", "double acc = 0;\n          if (e > 1) acc = 1;")),
           ("element_squared", ax.replace("w[e * n * n * n + k", "w[e * e * n * n * n + k")),
           ("missing_term", V_MULTIDIM[1].replace(" + dm[l][j] * ws[k * n * n + l * n + i]", ""))]
    for name, src in out:
        assert src not in (ax, V_MULTIDIM[1]), name
    return out
