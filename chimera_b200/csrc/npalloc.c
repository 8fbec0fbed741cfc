/* npalloc.c -- wires numpy's data allocator (NEP 49, PyDataMem_SetHandler) to the managed-memory functions of
 * libchimera_b200.so so that the arrays the reference's driver creates (np.zeros / np.empty / resize) are CUDA managed
 * memory and still OWN their data (the driver calls ndarray.resize on them, moduls/species.py:234-254).
 * Loaded with ctypes.PyDLL by chimera_b200/resident.py; no CUDA symbols here: the function pointers are passed in.
 * Blocks below `threshold` bytes stay with malloc (managed allocations are slow to create and 64 KB-granular). */
#define NPY_NO_DEPRECATED_API NPY_1_22_API_VERSION
#define NPY_TARGET_VERSION NPY_1_22_API_VERSION
#include <Python.h>
#include <numpy/arrayobject.h>
#include <stdlib.h>
#include <string.h>

typedef void* (*mm_alloc_t)(size_t, int);
typedef void* (*mm_realloc_t)(void*, size_t);
typedef void (*mm_free_t)(void*);
typedef int (*mm_owns_t)(const void*, size_t*);

static mm_alloc_t mm_alloc;
static mm_realloc_t mm_realloc;
static mm_free_t mm_free;
static mm_owns_t mm_owns;
static size_t g_threshold = 1 << 20;
static PyObject* g_prev = NULL;
static long long g_nmanaged = 0, g_bytes = 0;

static void* h_malloc(void* ctx, size_t size) {
  (void)ctx;
  if (size >= g_threshold) {
    void* p = mm_alloc(size, 0);
    if (p) { g_nmanaged++; g_bytes += (long long)size; return p; }
  }
  return malloc(size ? size : 1);
}
static void* h_calloc(void* ctx, size_t nelem, size_t elsize) {
  (void)ctx;
  const size_t size = nelem * elsize;
  if (size >= g_threshold) {
    void* p = mm_alloc(size, 1);
    if (p) { g_nmanaged++; g_bytes += (long long)size; return p; }
  }
  return calloc(nelem ? nelem : 1, elsize ? elsize : 1);
}
static void h_free(void* ctx, void* ptr, size_t size) {
  (void)ctx; (void)size;
  if (!ptr) return;
  if (mm_owns(ptr, NULL)) mm_free(ptr);
  else free(ptr);
}
static void* h_realloc(void* ctx, void* ptr, size_t new_size) {
  size_t old = 0;
  if (ptr && mm_owns(ptr, &old)) {
    if (new_size >= g_threshold) return mm_realloc(ptr, new_size);
    void* q = malloc(new_size ? new_size : 1);  /* shrinks below the threshold: back to malloc */
    if (!q) return NULL;
    memcpy(q, ptr, old < new_size ? old : new_size);
    mm_free(ptr);
    return q;
  }
  if (new_size >= g_threshold) {  /* grows over the threshold: move into managed memory */
    void* q = h_malloc(ctx, new_size);
    if (!q) return NULL;
    /* the old size is not passed in: realloc in place first, so that the copy length is known */
    void* t = realloc(ptr, new_size);
    if (!t) { h_free(ctx, q, new_size); return NULL; }
    memcpy(q, t, new_size);
    free(t);
    return q;
  }
  return realloc(ptr, new_size ? new_size : 1);
}

static PyDataMem_Handler g_handler = {"chimera_b200_managed", 1, {NULL, h_malloc, h_calloc, h_realloc, h_free}};

static int ensure_numpy(void) {
  if (PyArray_API == NULL) {
    if (_import_array() < 0) return -1;
  }
  return 0;
}

/* returns 0 on success; call with the GIL held (ctypes.PyDLL) */
int chb_npalloc_install(void* alloc_fn, void* realloc_fn, void* free_fn, void* owns_fn, size_t threshold) {
  if (ensure_numpy() < 0) return -1;
  if (g_prev) return 0; /* already installed */
  mm_alloc = (mm_alloc_t)alloc_fn; mm_realloc = (mm_realloc_t)realloc_fn; mm_free = (mm_free_t)free_fn; mm_owns = (mm_owns_t)owns_fn;
  g_threshold = threshold;
  PyObject* cap = PyCapsule_New(&g_handler, "mem_handler", NULL);
  if (!cap) return -2;
  g_prev = PyDataMem_SetHandler(cap);
  Py_DECREF(cap);
  return g_prev ? 0 : -3;
}
int chb_npalloc_uninstall(void) {
  if (!g_prev) return 0;
  PyObject* old = PyDataMem_SetHandler(g_prev);
  Py_XDECREF(old);
  Py_DECREF(g_prev);
  g_prev = NULL;
  return 0;
}
long long chb_npalloc_count(void) { return g_nmanaged; }
long long chb_npalloc_bytes(void) { return g_bytes; }
