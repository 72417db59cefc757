/* Evaluator for launch-size expressions.
 *
 * The reference keeps the kernel's grid/block sizes as SymEngine expressions and re-evaluates them whenever an
 * integer argument changes (reference src/symengine.c:13-68, src/loopy.c:406-449).  The expressions a loop nest
 * can produce are integer arithmetic over the kernel's integer arguments, so this file is a ~100-line recursive
 * descent evaluator instead of a computer-algebra dependency.  Grammar (emitted by nomp_bridge/emit_cuda.py):
 *     expr   := term (('+' | '-') term)*
 *     term   := factor (('*' | '/' | '%') factor)*          '/' truncates like C; divisor 0 is an error
 *     factor := INT | NAME | '(' expr ')' | ('+'|'-') factor | ('min'|'max') '(' expr ',' expr ')'
 */
#include <ctype.h>

#include "nomp-impl.h"

typedef struct {
  const char *p;
  const char *const *names;
  const long *values;
  unsigned n;
  int failed;
} gx_t;

static long gx_expr(gx_t *s);

static void gx_skip(gx_t *s) {
  while (isspace((unsigned char)*s->p)) s->p++;
}

static int gx_accept(gx_t *s, char ch) {
  gx_skip(s);
  if (*s->p != ch) return 0;
  s->p++;
  return 1;
}

static long gx_factor(gx_t *s) {
  gx_skip(s);
  if (gx_accept(s, '(')) {
    long v = gx_expr(s);
    if (!gx_accept(s, ')')) s->failed = 1;
    return v;
  }
  if (gx_accept(s, '-')) return -gx_factor(s);
  if (gx_accept(s, '+')) return gx_factor(s);
  if (isdigit((unsigned char)*s->p)) {
    char *end;
    long v = strtol(s->p, &end, 10);
    s->p = end;
    return v;
  }
  if (isalpha((unsigned char)*s->p) || *s->p == '_') {
    const char *b = s->p;
    while (isalnum((unsigned char)*s->p) || *s->p == '_') s->p++;
    size_t len = (size_t)(s->p - b);
    if ((len == 3) && (!strncmp(b, "min", 3) || !strncmp(b, "max", 3))) {
      const char *save = s->p;
      if (gx_accept(s, '(')) {
        int is_min = b[1] == 'i';
        long a = gx_expr(s);
        if (!gx_accept(s, ',')) s->failed = 1;
        long c = gx_expr(s);
        if (!gx_accept(s, ')')) s->failed = 1;
        return is_min ? (a < c ? a : c) : (a > c ? a : c);
      }
      s->p = save;
    }
    for (unsigned i = 0; i < s->n; i++) {
      if (strlen(s->names[i]) == len && !strncmp(s->names[i], b, len)) return s->values[i];
    }
  }
  s->failed = 1;
  return 0;
}

static long gx_term(gx_t *s) {
  long v = gx_factor(s);
  for (;;) {
    gx_skip(s);
    char op = *s->p;
    if (op != '*' && op != '/' && op != '%') return v;
    s->p++;
    long r = gx_factor(s);
    if (op == '*') v *= r;
    else if (r == 0) s->failed = 1;
    else if (op == '/') v /= r;
    else v %= r;
    if (s->failed) return 0;
  }
}

static long gx_expr(gx_t *s) {
  long v = gx_term(s);
  for (;;) {
    gx_skip(s);
    char op = *s->p;
    if (op != '+' && op != '-') return v;
    s->p++;
    long r = gx_term(s);
    v = op == '+' ? v + r : v - r;
    if (s->failed) return 0;
  }
}

/* 0 on success; 1 if the expression is malformed or names an unknown variable. */
int nomp_gridexpr_eval(const char *expr, const char *const *names, const long *values, unsigned n, long *result) {
  gx_t s = {expr, names, values, n, 0};
  long v = gx_expr(&s);
  gx_skip(&s);
  if (s.failed || *s.p != '\0') return 1;
  *result = v;
  return 0;
}
