/* libnomp core runtime: configuration, the host-range -> device-buffer table, the program table and the five
 * public entry points on the hot path (nomp_update, nomp_jit, nomp_run, nomp_sync, nomp_finalize).
 *
 * Behavioural contract = the reference's core (reference src/nomp.c:8-134 configuration, :258-365 mappings,
 * :367-571 jit, :603-656 run, :668-725 sync/finalize) including the quirks its tests pin.
 * What is different underneath:
 *   - mappings live in a hash table keyed by host pointer (the reference scans a list on every nomp_run argument,
 *     src/nomp.c:273-293, FIXME at :274);
 *   - launch sizes are strings evaluated by src/gridexpr.c, only when an integer argument changed, with no
 *     allocation on the nomp_run path (reference: 3 SymEngine objects per integer argument per run);
 *   - kernels receive the device address of host element 0 (bptr - idx0 * usize), so a loop over a sub-range
 *     mapping indexes the same elements on host and device;
 *   - the reduce clause is finished on the device (src/reduction.c -> backend).
 */
#include <ctype.h>

#include "nomp-aux.h"
#include "nomp-impl.h"
#include "nomp-loopy.h"

static nomp_backend_t nomp;
static nomp_config_t config; /* of the last successful nomp_init() */
static int initialized = 0;
static int device_reductions = 0; /* nomp_b200_device_reductions(): reduce results stay in mapped variables */

/* ============================================================================================================== */
/* configuration                                                                                                  */
/* ============================================================================================================== */
static void copy_bounded(char *dst, const char *src, size_t cap) {
  strncpy(dst, src, cap);
  dst[cap] = '\0';
}

static int parse_command_line(nomp_config_t *cfg, int argc, const char **argv) {
  if (argc <= 1 || argv == NULL) return 0;
  for (int i = 0; i < argc; i++) {
    const char *key = argv[i];
    if (key == NULL || strncmp(key, "--nomp", 6) != 0) continue; /* not ours: skip (reference src/nomp.c:47-50) */
    if (i + 1 >= argc || argv[i + 1] == NULL)
      return nomp_log(NOMP_USER_INPUT_IS_INVALID, NOMP_ERROR, "Missing argument value after: %s.", key);
    const char *val = argv[++i];
    if (!strcmp(key, "--nomp-install-dir")) copy_bounded(cfg->install_dir, val, PATH_MAX);
    else if (!strcmp(key, "--nomp-backend")) copy_bounded(cfg->backend, val, NOMP_MAX_BUFFER_SIZE);
    else if (!strcmp(key, "--nomp-platform")) cfg->platform = nomp_str_toui(val, NOMP_MAX_BUFFER_SIZE);
    else if (!strcmp(key, "--nomp-device")) cfg->device = nomp_str_toui(val, NOMP_MAX_BUFFER_SIZE);
    else if (!strcmp(key, "--nomp-verbose")) cfg->verbose = nomp_str_toui(val, NOMP_MAX_BUFFER_SIZE);
    else if (!strcmp(key, "--nomp-profile")) cfg->profile = nomp_str_toui(val, NOMP_MAX_BUFFER_SIZE);
    else if (!strcmp(key, "--nomp-scripts-dir")) copy_bounded(cfg->scripts_dir, val, PATH_MAX);
    else if (!strcmp(key, "--nomp-annotations-script")) copy_bounded(cfg->annotations_script, val, NOMP_MAX_BUFFER_SIZE);
    else nomp_log(NOMP_SUCCESS, NOMP_WARNING, "Unknown command line argument: %s.", key);
  }
  return 0;
}

static void apply_environment(nomp_config_t *cfg) {
  const char *v; /* the environment wins over the command line (reference src/nomp.c:106-107) */
  if ((v = getenv("NOMP_INSTALL_DIR"))) copy_bounded(cfg->install_dir, v, PATH_MAX);
  if ((v = getenv("NOMP_BACKEND"))) copy_bounded(cfg->backend, v, NOMP_MAX_BUFFER_SIZE);
  if ((v = getenv("NOMP_PLATFORM"))) cfg->platform = nomp_str_toui(v, NOMP_MAX_BUFFER_SIZE);
  if ((v = getenv("NOMP_DEVICE"))) cfg->device = nomp_str_toui(v, NOMP_MAX_BUFFER_SIZE);
  if ((v = getenv("NOMP_VERBOSE"))) cfg->verbose = nomp_str_toui(v, NOMP_MAX_BUFFER_SIZE);
  if ((v = getenv("NOMP_PROFILE"))) cfg->profile = nomp_str_toui(v, NOMP_MAX_BUFFER_SIZE);
  if ((v = getenv("NOMP_SCRIPTS_DIR"))) copy_bounded(cfg->scripts_dir, v, PATH_MAX);
}

static int missing_setting(const char *env, const char *flag) {
  return nomp_log(NOMP_USER_INPUT_IS_INVALID, NOMP_ERROR,
                  "%s is missing or invalid. Set it with %s command line argument or %s environment variable.", env,
                  flag, env);
}

static int load_config(nomp_config_t *cfg, int argc, const char **argv) {
  memset(cfg, 0, sizeof(*cfg));
  cfg->verbose = NOMP_DEFAULT_VERBOSE, cfg->profile = NOMP_DEFAULT_PROFILE;
  cfg->device = NOMP_DEFAULT_DEVICE, cfg->platform = NOMP_DEFAULT_PLATFORM;
  nomp_check(parse_command_line(cfg, argc, argv));
  apply_environment(cfg);
  for (char *p = cfg->backend; *p; p++) *p = (char)tolower((unsigned char)*p);

  if (cfg->install_dir[0] == '\0') return missing_setting("NOMP_INSTALL_DIR", "--nomp-install-dir");
  if (cfg->backend[0] == '\0') return missing_setting("NOMP_BACKEND", "--nomp-backend");
  if (cfg->verbose < 0) return missing_setting("NOMP_VERBOSE", "--nomp-verbose");
  if (cfg->profile < 0) return missing_setting("NOMP_PROFILE", "--nomp-profile");
  if (cfg->device < 0) return missing_setting("NOMP_DEVICE", "--nomp-device");
  if (cfg->platform < 0) return missing_setting("NOMP_PLATFORM", "--nomp-platform");
  return 0;
}

/* ============================================================================================================== */
/* init                                                                                                           */
/* ============================================================================================================== */
static int init_backend(nomp_backend_t *bnd, const nomp_config_t *cfg) {
  bnd->py_context = nomp_py_dict_new();
  nomp_py_dict_set_str(bnd->py_context, "backend::name", cfg->backend);
  if (!strcmp(cfg->backend, "cuda")) return cuda_init(bnd, cfg->platform, cfg->device);
  /* CUDA only: no HIP / OpenCL dispatch in this implementation */
  return nomp_log(NOMP_USER_INPUT_IS_INVALID, NOMP_ERROR, "Invalid backend: %s.", cfg->backend);
}

static int allocate_scratch(nomp_backend_t *bnd) {
  nomp_mem_t *m = &bnd->scratch;
  memset(m, 0, sizeof(*m));
  m->idx0 = 0, m->idx1 = NOMP_MAX_SCRATCH_SIZE, m->usize = sizeof(double);
  return bnd->update(bnd, m, NOMP_ALLOC, m->idx0, m->idx1, m->usize);
}

static void release_partial_init(void) {
  nomp_py_decref(&nomp.py_annotate);
  nomp_py_decref(&nomp.py_context);
  if (nomp.bptr && nomp.finalize) nomp.finalize(&nomp);
  memset(&nomp, 0, sizeof(nomp));
}

NOMP_EXPORT int nomp_init(int argc, const char **argv) {
  nomp_log_set_verbose(NOMP_DEFAULT_VERBOSE);
  if (initialized) return nomp_log(NOMP_INITIALIZE_FAILURE, NOMP_ERROR, "libnomp is already initialized.");

  nomp_config_t cfg;
  nomp_check(load_config(&cfg, argc, argv));
  nomp_check(nomp_py_init(&cfg));
  nomp_check(nomp_profile_set_level(cfg.profile));
  nomp_check(nomp_log_set_verbose((unsigned)cfg.verbose));

  memset(&nomp, 0, sizeof(nomp));
  int err = nomp_py_set_annotate_func(&nomp.py_annotate, cfg.annotations_script);
  if (!err) err = init_backend(&nomp, &cfg);
  if (!err) err = allocate_scratch(&nomp);
  if (err) {
    release_partial_init();
    return err;
  }
  config = cfg;
  nomp_jit_cache_reset();
  initialized = 1;
  device_reductions = 0;
  return 0;
}

/* ============================================================================================================== */
/* host range -> device buffer table                                                                              */
/* ============================================================================================================== */
typedef struct mem_node {
  nomp_mem_t m;
  struct mem_node *next;
} mem_node_t;

static mem_node_t **buckets = NULL;
static size_t n_buckets = 0, n_mems = 0;

static size_t bucket_of(const void *p, size_t nb) {
  uint64_t x = (uint64_t)(uintptr_t)p;
  x ^= x >> 33, x *= 0xff51afd7ed558ccdULL, x ^= x >> 33;
  return (size_t)(x & (nb - 1));
}

static void table_grow(void) {
  size_t nb = n_buckets ? 2 * n_buckets : 64;
  mem_node_t **fresh = nomp_calloc(mem_node_t *, nb);
  for (size_t b = 0; b < n_buckets; b++) {
    for (mem_node_t *n = buckets[b], *next; n; n = next) {
      next = n->next;
      size_t t = bucket_of(n->m.hptr, nb);
      n->next = fresh[t], fresh[t] = n;
    }
  }
  free(buckets);
  buckets = fresh, n_buckets = nb;
}

/* first mapping whose host pointer is exactly `hptr` (what a kernel argument is matched against) */
nomp_mem_t *nomp_lookup_mem(const void *hptr) {
  if (n_buckets == 0 || hptr == NULL) return NULL;
  for (mem_node_t *n = buckets[bucket_of(hptr, n_buckets)]; n; n = n->next)
    if (n->m.hptr == hptr) return &n->m;
  return NULL;
}

/* mapping of `hptr` that covers bytes [idx0*usize, idx1*usize) (reference src/nomp.c:282-293: pointer AND byte range) */
static mem_node_t *lookup_range(const void *hptr, size_t idx0, size_t idx1, size_t usize) {
  if (n_buckets == 0) return NULL;
  for (mem_node_t *n = buckets[bucket_of(hptr, n_buckets)]; n; n = n->next) {
    if (n->m.hptr == hptr && n->m.idx0 * n->m.usize <= idx0 * usize && n->m.idx1 * n->m.usize >= idx1 * usize)
      return n;
  }
  return NULL;
}

static void table_remove(mem_node_t *node) {
  mem_node_t **link = &buckets[bucket_of(node->m.hptr, n_buckets)];
  while (*link && *link != node) link = &(*link)->next;
  if (*link) *link = node->next;
  free(node);
  n_mems--;
}

/* Versions of device images come from one counter that never restarts: a mapping created after NOMP_FREE at the same
 * host and device addresses cannot repeat a (pointer, version) pair that a cache (the staged D of the Ax family) holds. */
unsigned long nomp_next_version(void) {
  static unsigned long counter = 0;
  return ++counter;
}

NOMP_EXPORT int nomp_update(void *ptr, size_t idx0, size_t idx1, size_t unit_size, nomp_map_direction_t op) {
  if (!initialized) return nomp_log(NOMP_INITIALIZE_FAILURE, NOMP_ERROR, "libnomp is not initialized.");
  mem_node_t *node = lookup_range(ptr, idx0, idx1, unit_size);
  int created = 0;
  if (node == NULL) {
    if (op == NOMP_FROM || op == NOMP_FREE) {
      return nomp_log(NOMP_USER_MAP_OP_IS_INVALID, NOMP_ERROR,
                      "NOMP_FREE or NOMP_FROM can only be called on a pointer which is already on the device.");
    }
    op |= NOMP_ALLOC; /* NOMP_TO on a new range implies allocation */
    node = nomp_calloc(mem_node_t, 1);
    node->m.idx0 = idx0, node->m.idx1 = idx1, node->m.usize = unit_size, node->m.hptr = ptr;
    node->m.version = nomp_next_version();
    created = 1;
  }

  int err = nomp.update(&nomp, &node->m, op, idx0, idx1, unit_size);
  if (err > 0) {
    if (created) free(node);
    return err;
  }
  if (created) {
    if (n_mems + 1 > n_buckets / 2) table_grow();
    size_t b = bucket_of(ptr, n_buckets);
    node->next = buckets[b], buckets[b] = node, n_mems++;
  } else if (node->m.bptr == NULL) {
    table_remove(node); /* the backend released the buffer */
  }
  return 0;
}

/* include/nomp-b200.h: asynchronous NOMP_TO / NOMP_FROM on an EXISTING mapping; completion at nomp_sync(). */
NOMP_EXPORT int nomp_b200_update_async(void *ptr, size_t idx0, size_t idx1, size_t unit_size, nomp_map_direction_t op) {
  if (!initialized) return nomp_log(NOMP_INITIALIZE_FAILURE, NOMP_ERROR, "libnomp is not initialized.");
  if (op != NOMP_TO && op != NOMP_FROM)
    return nomp_log(NOMP_USER_MAP_OP_IS_INVALID, NOMP_ERROR, "nomp_b200_update_async supports NOMP_TO and NOMP_FROM only.");
  mem_node_t *node = lookup_range(ptr, idx0, idx1, unit_size);
  if (node == NULL)
    return nomp_log(NOMP_USER_MAP_OP_IS_INVALID, NOMP_ERROR,
                    "nomp_b200_update_async can only be called on a range which is already on the device.");
  return nomp_cuda_update_async(&nomp, &node->m, op, idx0, idx1, unit_size);
}

/* include/nomp-b200.h: leave the result of a reduce clause in the device copy of its variable when that variable is
 * mapped, and do not wait for it.  Off by default: the reference writes the host variable before nomp_run returns. */
NOMP_EXPORT int nomp_b200_device_reductions(int enable) {
  const int before = device_reductions;
  if (enable >= 0) device_reductions = enable != 0;
  return before;
}

NOMP_EXPORT void *nomp_b200_device_ptr(void *hptr) {
  nomp_mem_t *m = nomp_lookup_mem(hptr);
  return m ? (char *)m->bptr - m->idx0 * m->usize : NULL;
}

/* ============================================================================================================== */
/* programs                                                                                                       */
/* ============================================================================================================== */
static nomp_prog_t **progs = NULL;
static unsigned progs_n = 0, progs_max = 0;

static void free_prog(nomp_prog_t *prg) {
  if (prg == NULL) return;
  if (prg->bptr && nomp.knl_free) nomp.knl_free(prg);
  nomp_py_decref(&prg->py_dict);
  for (int d = 0; d < 3; d++) free(prg->sym_global[d]), free(prg->sym_local[d]);
  free(prg->info);
  free(prg->args);
  free(prg);
}

/* one (name, size, type[, value]) group per kernel argument; NOMP_JIT arguments go to py_dict instead */
static nomp_prog_t *collect_args(unsigned nargs, va_list ap) {
  nomp_prog_t *prg = nomp_calloc(nomp_prog_t, 1);
  prg->args = nomp_calloc(nomp_arg_t, nargs + 1);
  prg->reduction_index = -1;
  prg->py_dict = nomp_py_dict_new();
  for (unsigned i = 0; i < nargs; i++) {
    const char *name = va_arg(ap, const char *);
    const size_t size = va_arg(ap, size_t);
    int type = va_arg(ap, int);
    if (type & NOMP_JIT) {
      type &= ~NOMP_JIT;
      const void *value = va_arg(ap, void *);
      if (value == NULL) continue;
      switch (type) {
      case NOMP_INT:
        nomp_py_dict_set_long(prg->py_dict, name, size == 8 ? *(const long *)value : (long)*(const int *)value);
        break;
      case NOMP_UINT:
        nomp_py_dict_set_long(prg->py_dict, name,
                              size == 8 ? (long)*(const unsigned long *)value : (long)*(const unsigned *)value);
        break;
      case NOMP_FLOAT:
        nomp_py_dict_set_double(prg->py_dict, name, size == 4 ? (double)*(const float *)value : *(const double *)value);
        break;
      default: break;
      }
      continue;
    }
    nomp_arg_t *a = &prg->args[prg->nargs++];
    copy_bounded(a->name, name, NOMP_MAX_BUFFER_SIZE);
    a->size = size, a->type = (nomp_arg_type_t)type;
  }
  return prg;
}

/* with_python = 0: only the C side (validation, reduction bookkeeping), for programs served from the JIT cache */
static int apply_clauses(PyObject **knl, nomp_prog_t *prg, const char **clauses, const char **reduce_op,
                         int with_python) {
  *reduce_op = NULL;
  for (unsigned i = 0; clauses && clauses[i]; i += 3) {
    const char *kind = clauses[i];
    if (!strcmp(kind, "transform")) {
      const char *file = clauses[i + 1], *function = clauses[i + 2];
      if (!with_python) continue;
      nomp_check(nomp_py_check_module(file, function));
      nomp_check(nomp_py_transform(knl, file, function, nomp.py_context));
    } else if (!strcmp(kind, "reduce")) {
      const char *var = clauses[i + 1], *op = clauses[i + 2];
      if (var == NULL || op == NULL)
        return nomp_log(NOMP_USER_INPUT_IS_INVALID, NOMP_ERROR, "The reduce clause needs a variable and an operator.");
      for (unsigned j = 0; j < prg->nargs; j++) {
        if (strncmp(prg->args[j].name, var, NOMP_MAX_BUFFER_SIZE)) continue;
        /* the accumulator is declared with its scalar type and becomes a pointer argument (reference src/nomp.c:388-397) */
        prg->reduction_type = prg->args[j].type, prg->reduction_size = (int)prg->args[j].size;
        prg->reduction_index = (int)j, prg->args[j].type = NOMP_PTR;
        break;
      }
      if (prg->reduction_index < 0)
        return nomp_log(NOMP_USER_INPUT_IS_INVALID, NOMP_ERROR,
                        "Reduction variable \"%s\" is not an argument of the kernel.", var);
      if (!strcmp(op, "+")) prg->reduction_op = NOMP_SUM;
      else if (!strcmp(op, "*")) prg->reduction_op = NOMP_PROD;
      else if (!strcmp(op, "min")) prg->reduction_op = NOMP_MIN;
      else if (!strcmp(op, "max")) prg->reduction_op = NOMP_MAX;
      else
        return nomp_log(NOMP_USER_INPUT_IS_INVALID, NOMP_ERROR,
                        "Reduction operator \"%s\" is not one of \"+\", \"*\", \"min\", \"max\".", op);
      *reduce_op = op;
    } else if (!strcmp(kind, "annotate")) {
      if (!with_python) continue;
      PyObject *annotations = nomp_py_dict_new();
      nomp_py_dict_set_str(annotations, clauses[i + 1], clauses[i + 2] ? clauses[i + 2] : "");
      int err = nomp_py_annotate(knl, nomp.py_annotate, annotations, nomp.py_context);
      nomp_py_decref(&annotations);
      if (err) return err;
    } else {
      return nomp_log(NOMP_USER_INPUT_IS_INVALID, NOMP_ERROR,
                      "Clause \"%s\" passed into nomp_jit is not a valid clause.", kind);
    }
  }
  return 0;
}

/* ---- on-disk cache of what the bridge produces for one nomp_jit() (src/jitcache.c) ------------------------------ */
/* SHA-256 of the bridge's own Python text: a new bridge version never reads entries of an old one. */
static const char *bridge_digest(void) {
  static char digest[65], of_dir[PATH_MAX + 1];
  if (digest[0] && !strcmp(of_dir, config.install_dir)) return digest[0] == '-' ? NULL : digest;
  snprintf(of_dir, sizeof(of_dir), "%s", config.install_dir);
  nomp_sha256_t c;
  nomp_sha256_init(&c);
  char dir[PATH_MAX + 64];
  snprintf(dir, sizeof(dir), "%s/python/nomp_bridge", config.install_dir);
  int err = nomp_sha256_dir(&c, dir, ".py");
  snprintf(dir, sizeof(dir), "%s/python/loopy", config.install_dir);
  err |= nomp_sha256_dir(&c, dir, ".py");
  if (err) {
    strcpy(digest, "-");
    return NULL;
  }
  nomp_sha256_hex(&c, digest);
  return digest;
}

static int hash_script(nomp_sha256_t *c, const char *module) {
  char *path = NULL;
  if (module == NULL || nomp_py_module_file(&path, module)) return 1;
  const int err = nomp_sha256_file(c, path);
  free(path);
  return err;
}

/* Key of a program; non-zero when it cannot be cached (cache off, a script that cannot be located, ...). */
static int program_key(char hex[65], const nomp_prog_t *prg, const char *csrc, const char **clauses) {
  if (nomp_jit_cache_dir() == NULL) return 1;
  const char *bridge = bridge_digest();
  if (bridge == NULL) return 1;
  nomp_sha256_t c;
  nomp_sha256_init(&c);
  nomp_sha256_field(&c, "libnomp_b200 knl 1");
  nomp_sha256_field(&c, bridge);
  nomp_sha256_field(&c, csrc);
  int annotated = 0;
  for (unsigned i = 0; clauses && clauses[i]; i += 3) {
    nomp_sha256_field(&c, clauses[i]), nomp_sha256_field(&c, clauses[i + 1]), nomp_sha256_field(&c, clauses[i + 2]);
    if (!strcmp(clauses[i], "transform") && hash_script(&c, clauses[i + 1])) return 1;
    annotated |= !strcmp(clauses[i], "annotate");
  }
  if (annotated && nomp.py_annotate && hash_script(&c, config.annotations_script)) return 1;
  for (unsigned i = 0; i < prg->nargs; i++) {
    char buf[64];
    snprintf(buf, sizeof(buf), "%zu:%d", prg->args[i].size, (int)prg->args[i].type);
    nomp_sha256_field(&c, prg->args[i].name), nomp_sha256_field(&c, buf);
  }
  char *jit_values = NULL, *context = NULL;
  int err = nomp_py_repr(&jit_values, prg->py_dict) || nomp_py_repr(&context, nomp.py_context);
  if (!err) nomp_sha256_field(&c, jit_values), nomp_sha256_field(&c, context);
  free(jit_values), free(context);
  if (!err) nomp_sha256_hex(&c, hex);
  return err;
}

#define KNL_ENTRY_MAGIC "NOMPJIT1\n"

static void store_program(const char *hex, const nomp_prog_t *prg, const char *name, const char *src) {
  size_t cap = strlen(KNL_ENTRY_MAGIC) + strlen(name) + strlen(src) + 16;
  for (int d = 0; d < 3; d++) cap += strlen(prg->sym_global[d]) + strlen(prg->sym_local[d]);
  char *buf = nomp_calloc(char, cap);
  const int n = snprintf(buf, cap, KNL_ENTRY_MAGIC "%s\n%s\n%s\n%s\n%s\n%s\n%s\n%s", name, prg->sym_global[0],
                         prg->sym_global[1], prg->sym_global[2], prg->sym_local[0], prg->sym_local[1], prg->sym_local[2], src);
  if (n > 0 && (size_t)n < cap) nomp_jit_cache_store(hex, "knl", buf, (size_t)n);
  free(buf);
}

/* name, six launch-size expressions (one per line), then the generated source up to the end of the file */
static int load_program(const char *hex, nomp_prog_t *prg, char **name, char **src) {
  char *data = NULL;
  size_t size = 0;
  if (nomp_jit_cache_load(hex, "knl", &data, &size)) return 1;
  const size_t lm = strlen(KNL_ENTRY_MAGIC);
  char *line[7], *p = data + lm;
  int ok = size > lm && !memcmp(data, KNL_ENTRY_MAGIC, lm) && strlen(data) == size;
  for (int i = 0; ok && i < 7; i++) {
    char *eol = strchr(p, '\n');
    if (!(ok = eol != NULL)) break;
    *eol = '\0', line[i] = p, p = eol + 1;
  }
  ok = ok && !strncmp(p, "//!nomp ", 8);
  if (ok) {
    *name = strdup(line[0]), *src = strdup(p);
    for (int d = 0; d < 3; d++) {
      free(prg->sym_global[d]), free(prg->sym_local[d]);
      prg->sym_global[d] = strdup(line[1 + d]), prg->sym_local[d] = strdup(line[4 + d]);
    }
    prg->ndim = 3;
  }
  free(data);
  return !ok;
}

static void set_info(nomp_prog_t *prg, const char *src) {
  const char *eol = strchr(src, '\n');
  prg->info = strndup(src + 8, eol ? (size_t)(eol - src) - 8 : strlen(src) - 8);
}

static int build_program(nomp_prog_t *prg, const char *csrc, const char **clauses) {
  const char *reduce_op = NULL;
  char key[65], *name = NULL, *src = NULL;
  const int cacheable = !program_key(key, prg, csrc, clauses);
  if (cacheable && !load_program(key, prg, &name, &src)) {
    /* served from the cache: the interpreter is not entered, user scripts do not run */
    int err = apply_clauses(NULL, prg, clauses, &reduce_op, 0);
    if (!err) {
      set_info(prg, src);
      err = nomp.knl_build(&nomp, prg, src, name);
    }
    free(name), free(src);
    if (!err) nomp_jit_cache_count(NOMP_CACHE_KNL_HIT);
    return err;
  }

  PyObject *knl = NULL;
  nomp_check(nomp_py_c_to_loopy(&knl, csrc));

  int err = apply_clauses(&knl, prg, clauses, &reduce_op, 1);
  if (!err && prg->reduction_index >= 0)
    err = nomp_py_realize_reduction(&knl, prg->args[prg->reduction_index].name, reduce_op, nomp.py_context);
  if (!err && nomp_py_dict_size(prg->py_dict) > 0) err = nomp_py_fix_parameters(&knl, prg->py_dict);

  if (!err) err = nomp_py_get_knl_name_and_src(&name, &src, knl, nomp.py_context);
  if (!err) {
    set_info(prg, src);
    err = nomp.knl_build(&nomp, prg, src, name);
  }
  if (!err) err = nomp_py_get_grid_size(prg, knl, nomp.py_context);
  if (!err && cacheable) {
    store_program(key, prg, name, src);
    nomp_jit_cache_count(NOMP_CACHE_KNL_MISS);
  }
  free(name), free(src);
  nomp_py_decref(&knl);
  return err;
}

NOMP_EXPORT int nomp_jit(int *id, const char *csrc, const char **clauses, int nargs, ...) {
  if (*id >= 0) return 0; /* the caller's static id is the kernel cache (reference src/nomp.c:527) */
  if (!initialized) return nomp_log(NOMP_INITIALIZE_FAILURE, NOMP_ERROR, "libnomp is not initialized.");
  if (nargs < 0 || nargs > NOMP_MAX_KERNEL_ARGS_SIZE)
    return nomp_log(NOMP_USER_INPUT_IS_INVALID, NOMP_ERROR, "A kernel can have at most %d arguments.",
                    NOMP_MAX_KERNEL_ARGS_SIZE);

  va_list ap;
  va_start(ap, nargs);
  nomp_prog_t *prg = collect_args((unsigned)nargs, ap);
  va_end(ap);

  int err = build_program(prg, csrc, clauses);
  if (err) {
    free_prog(prg);
    return err;
  }
  if (progs_n == progs_max) {
    progs_max = progs_max ? 2 * progs_max : 16;
    progs = nomp_realloc(progs, nomp_prog_t *, progs_max);
  }
  progs[progs_n] = prg;
  *id = (int)progs_n++;
  return 0;
}

static int evaluate_launch_size(nomp_prog_t *prg) {
  const char *names[NOMP_MAX_KERNEL_ARGS_SIZE];
  long values[NOMP_MAX_KERNEL_ARGS_SIZE];
  unsigned n = 0;
  for (unsigned i = 0; i < prg->nargs; i++) {
    if (prg->args[i].type == NOMP_INT || prg->args[i].type == NOMP_UINT)
      names[n] = prg->args[i].name, values[n] = prg->int_values[i], n++;
  }
  for (int d = 0; d < 3; d++) {
    long g = 1, l = 1;
    if (nomp_gridexpr_eval(prg->sym_global[d] ? prg->sym_global[d] : "1", names, values, n, &g) ||
        nomp_gridexpr_eval(prg->sym_local[d] ? prg->sym_local[d] : "1", names, values, n, &l))
      return nomp_log(NOMP_LOOPY_GRIDSIZE_FAILURE, NOMP_ERROR, "Unable to evaluate grid sizes from loopy kernel.");
    prg->global[d] = g < 0 ? 0 : (size_t)g;
    prg->local[d] = l < 1 ? 1 : (size_t)l;
    prg->gws[d] = prg->global[d] * prg->local[d];
  }
  return 0;
}

NOMP_EXPORT int nomp_run(int id, ...) {
  if (id < 0 || (unsigned)id >= progs_n || progs[id] == NULL)
    return nomp_log(NOMP_USER_INPUT_IS_INVALID, NOMP_ERROR, "Kernel id %d passed to nomp_run is not valid.", id);

  nomp_prog_t *prg = progs[id];
  nomp_arg_t *args = prg->args;
  int eval_grid = !prg->int_valid; /* first run: sizes were never evaluated (the reference launches with zero
                                      dimensions when a kernel has no integer argument, src/nomp.c:610-649) */
  va_list ap;
  va_start(ap, id);
  for (unsigned i = 0; i < prg->nargs; i++) {
    void *p = va_arg(ap, void *);
    args[i].ptr = p, args[i].mem = NULL;
    switch (args[i].type) {
    case NOMP_INT:
    case NOMP_UINT: {
      long v;
      if (args[i].type == NOMP_INT) v = args[i].size == 8 ? *(long *)p : (long)*(int *)p;
      else v = args[i].size == 8 ? (long)*(unsigned long *)p : (long)*(unsigned *)p;
      if (v != prg->int_values[i]) eval_grid = 1;
      prg->int_values[i] = v;
      break;
    }
    case NOMP_PTR: {
      nomp_mem_t *m = nomp_lookup_mem(p);
      if (m == NULL) {
        if (prg->reduction_index == (int)i) { /* the accumulator is an unmapped host address */
          prg->reduction_ptr = p, prg->reduction_dev = NULL, args[i].ptr = NULL;
          break;
        }
        va_end(ap);
        return nomp_log(NOMP_USER_MAP_PTR_IS_INVALID, NOMP_ERROR, ERR_STR_USER_MAP_PTR_IS_INVALID, p);
      }
      args[i].mem = m;
      args[i].ptr = (char *)m->bptr - m->idx0 * m->usize; /* device address of host element 0 */
      if (prg->reduction_index == (int)i)
        prg->reduction_ptr = p, prg->reduction_dev = device_reductions ? args[i].ptr : NULL;
      break;
    }
    default: break; /* NOMP_FLOAT: the pointer to the scalar is passed through */
    }
  }
  va_end(ap);

  if (eval_grid) {
    nomp_check(evaluate_launch_size(prg));
    prg->int_valid = 1;
  }
  nomp_check(nomp.knl_run(&nomp, prg));
  for (unsigned i = 0; i < prg->nargs; i++) {
    if (args[i].mem && !args[i].is_const) ((nomp_mem_t *)args[i].mem)->version = nomp_next_version();
  }
  if (prg->reduction_index >= 0) nomp_check(nomp_device_side_reduction(&nomp, prg));
  return 0;
}

NOMP_EXPORT int nomp_sync(void) {
  if (!initialized) return nomp_log(NOMP_INITIALIZE_FAILURE, NOMP_ERROR, "libnomp is not initialized.");
  nomp_check(nomp.sync(&nomp));
  return nomp_gs_check(); /* a gather-scatter that gave up waiting for a peer rank reports it here */
}

NOMP_EXPORT const char *nomp_b200_prog_info(int id) {
  if (id < 0 || (unsigned)id >= progs_n || progs[id] == NULL) return NULL;
  return progs[id]->info;
}

/* ============================================================================================================== */
/* finalize                                                                                                       */
/* ============================================================================================================== */
static int finalize_impl(int interpreter) {
  if (!initialized) return NOMP_FINALIZE_FAILURE; /* raw code, no log entry (reference src/nomp.c:671) */

  nomp_gs_finalize();
  nomp_py_decref(&nomp.py_annotate);
  nomp_py_decref(&nomp.py_context);

  for (size_t b = 0; b < n_buckets; b++) {
    for (mem_node_t *n = buckets[b], *next; n; n = next) {
      next = n->next;
      if (n->m.bptr) nomp_check(nomp.update(&nomp, &n->m, NOMP_FREE, n->m.idx0, n->m.idx1, n->m.usize));
      free(n);
    }
    buckets[b] = NULL;
  }
  free(buckets);
  buckets = NULL, n_buckets = n_mems = 0;
  if (nomp.scratch.bptr)
    nomp_check(nomp.update(&nomp, &nomp.scratch, NOMP_FREE, nomp.scratch.idx0, nomp.scratch.idx1, nomp.scratch.usize));

  for (unsigned i = 0; i < progs_n; i++) free_prog(progs[i]), progs[i] = NULL;
  free(progs);
  progs = NULL, progs_n = progs_max = 0;

  nomp_check(nomp_py_finalize(interpreter));
  nomp_profile_finalize();
  nomp_log_finalize();

  initialized = nomp.finalize(&nomp);
  if (initialized) return NOMP_FINALIZE_FAILURE;
  memset(&nomp, 0, sizeof(nomp));
  return 0;
}

NOMP_EXPORT int nomp_finalize(void) { return finalize_impl(1); }
NOMP_EXPORT int nomp_finalize_excluding_interpreter(void) { return finalize_impl(0); }
