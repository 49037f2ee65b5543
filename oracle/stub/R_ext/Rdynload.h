/* Stand-in for <R_ext/Rdynload.h>; test infrastructure only (see ../Rinternals.h). */
#ifndef EDB200_STUB_RDYNLOAD_H
#define EDB200_STUB_RDYNLOAD_H
#ifdef __cplusplus
extern "C" {
#endif
typedef void *(*DL_FUNC)(void);
typedef struct { const char *name; DL_FUNC fun; int numArgs; } R_CallMethodDef;
typedef struct edb200_stub_dllinfo {
    const R_CallMethodDef *call_entries;
    int use_dynamic_symbols;
} DllInfo;
typedef int Rboolean;
#ifndef FALSE
#define FALSE 0
#endif
#ifndef TRUE
#define TRUE 1
#endif
int R_registerRoutines(DllInfo *dll, const void *c_entries, const R_CallMethodDef *call_entries,
                       const void *fortran_entries, const void *external_entries);
int R_useDynamicSymbols(DllInfo *dll, Rboolean value);
#ifdef __cplusplus
}
#endif
#endif
