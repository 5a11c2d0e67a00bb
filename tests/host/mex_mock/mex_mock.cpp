// mex_mock.cpp -- a stand-in for the part of MATLAB's mx/mex runtime that the gateways in subzero_b200/matlab/ use
// (test infrastructure; MATLAB is not installed in this image).  It implements the declarations of
// tests/host/mex_stub/mex.h with ordinary heap objects so that the UNMODIFIED gateway sources can be compiled,
// loaded and *called* from the test-suite: the same mexFunction(nlhs, plhs, nrhs, prhs) entry point MATLAB would call
// (private/mexclipper.cpp:83 is the reference's), with struct / double / char arguments built by the test.
//
// Semantics kept from MATLAB where the gateways depend on them:
//   * mexErrMsgIdAndTxt does not return (MATLAB longjmps out of the mex file; here a C++ exception unwinds to mm_call);
//   * arrays are column-major doubles, mxGetNumberOfElements = rows * cols, mxIsEmpty = no elements;
//   * mxGetField returns NULL for an absent field, mxGetScalar reads the first element;
//   * functions registered with mexAtExit run when the mex file is cleared (mm_clear).
// Not kept: memory management (arrays live until mm_free_all), every type but real double / 1x1 struct / char row.
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

struct mxArray_tag {
    enum Kind { DOUBLE, STRUCT, CHAR } kind = DOUBLE;
    size_t m = 0, n = 0;
    std::vector<double> d;                 // DOUBLE: column-major
    std::vector<std::string> names;        // STRUCT (1 x 1): field names ...
    std::vector<mxArray_tag*> fields;      // ... and values (NULL = the [] MATLAB leaves in an unset field)
    std::string s;                         // CHAR
};
typedef mxArray_tag mxArray;
typedef enum { mxREAL = 0, mxCOMPLEX = 1 } mxComplexity;

namespace {
struct MexError { std::string id, msg; };
std::vector<mxArray*> g_all;
std::vector<void (*)(void)> g_atexit;
mxArray* track(mxArray* a) { g_all.push_back(a); return a; }
}

extern "C" {
bool mxIsChar(const mxArray* a) { return a && a->kind == mxArray::CHAR; }
int mxGetString(const mxArray* a, char* buf, size_t cap)
{
    if (!a || a->kind != mxArray::CHAR || cap == 0) return 1;
    const size_t k = a->s.size() < cap - 1 ? a->s.size() : cap - 1;
    std::memcpy(buf, a->s.data(), k); buf[k] = 0;
    return a->s.size() >= cap;             // MATLAB: 1 when the string was truncated
}
bool mxIsStruct(const mxArray* a) { return a && a->kind == mxArray::STRUCT; }
bool mxIsDouble(const mxArray* a) { return a && a->kind == mxArray::DOUBLE; }
bool mxIsComplex(const mxArray*) { return false; }
bool mxIsEmpty(const mxArray* a) { return !a || (a->kind == mxArray::CHAR ? a->s.empty() : a->m * a->n == 0); }
mxArray* mxGetField(const mxArray* a, size_t idx, const char* name)
{
    if (!a || a->kind != mxArray::STRUCT || idx != 0) return nullptr;
    for (size_t k = 0; k < a->names.size(); ++k) if (a->names[k] == name) return a->fields[k];
    return nullptr;
}
size_t mxGetNumberOfElements(const mxArray* a) { return !a ? 0 : a->kind == mxArray::CHAR ? a->s.size() : a->m * a->n; }
double* mxGetPr(const mxArray* a) { return a && a->kind == mxArray::DOUBLE ? const_cast<double*>(a->d.data()) : nullptr; }
double mxGetScalar(const mxArray* a)
{
    if (!a || a->kind != mxArray::DOUBLE || a->d.empty()) throw MexError{"MATLAB:mxGetScalar", "mxGetScalar of an empty or non-numeric array"};
    return a->d[0];
}
mxArray* mxCreateDoubleMatrix(size_t m, size_t n, mxComplexity)
{
    mxArray* a = track(new mxArray); a->kind = mxArray::DOUBLE; a->m = m; a->n = n; a->d.assign(m * n, 0.0); return a;
}
mxArray* mxCreateDoubleScalar(double v) { mxArray* a = mxCreateDoubleMatrix(1, 1, mxREAL); a->d[0] = v; return a; }
mxArray* mxCreateStructMatrix(size_t m, size_t n, int nfields, const char** names)
{
    mxArray* a = track(new mxArray); a->kind = mxArray::STRUCT; a->m = m; a->n = n;
    for (int k = 0; k < nfields; ++k) { a->names.push_back(names[k]); a->fields.push_back(nullptr); }
    return a;
}
void mxSetFieldByNumber(mxArray* a, size_t idx, int k, mxArray* v)
{
    if (!a || a->kind != mxArray::STRUCT || idx != 0 || k < 0 || (size_t)k >= a->fields.size()) throw MexError{"MATLAB:mxSetFieldByNumber", "bad struct, index or field number"};
    a->fields[(size_t)k] = v;
}
void mexErrMsgIdAndTxt(const char* id, const char* fmt, ...)
{
    char buf[2048];
    va_list ap; va_start(ap, fmt); std::vsnprintf(buf, sizeof buf, fmt, ap); va_end(ap);
    throw MexError{id ? id : "", buf};
}
int mexAtExit(void (*fn)(void)) { g_atexit.push_back(fn); return 0; }

// ---- the test's side: build arguments, call a gateway, read results (ctypes, tests/test_mex_gateway.py)
mxArray* mm_double(size_t m, size_t n, const double* data)
{
    mxArray* a = mxCreateDoubleMatrix(m, n, mxREAL);
    if (data && m * n != 0) std::memcpy(a->d.data(), data, m * n * sizeof(double));
    return a;
}
mxArray* mm_string(const char* s) { mxArray* a = track(new mxArray); a->kind = mxArray::CHAR; a->s = s; a->m = 1; a->n = a->s.size(); return a; }
mxArray* mm_struct() { mxArray* a = track(new mxArray); a->kind = mxArray::STRUCT; a->m = a->n = 1; return a; }
void mm_set(mxArray* st, const char* name, mxArray* v)
{
    for (size_t k = 0; k < st->names.size(); ++k) if (st->names[k] == name) { st->fields[k] = v; return; }
    st->names.push_back(name); st->fields.push_back(v);
}
int mm_kind(const mxArray* a) { return a ? (int)a->kind : -1; }
size_t mm_rows(const mxArray* a) { return a ? a->m : 0; }
size_t mm_cols(const mxArray* a) { return a ? a->n : 0; }
int mm_nfields(const mxArray* a) { return a && a->kind == mxArray::STRUCT ? (int)a->names.size() : 0; }
const char* mm_field_name(const mxArray* a, int k) { return a->names[(size_t)k].c_str(); }

typedef void (*MexFunction)(int, mxArray**, int, const mxArray**);
// Calls a gateway the way MATLAB does.  Returns 0, or 1 when the gateway raised an error: `err` then holds "id|message".
int mm_call(MexFunction fn, int nlhs, mxArray** plhs, int nrhs, const mxArray** prhs, char* err, size_t errcap)
{
    try { fn(nlhs, plhs, nrhs, prhs); }
    catch (const MexError& e) { std::snprintf(err, errcap, "%s|%s", e.id.c_str(), e.msg.c_str()); return 1; }
    return 0;
}
// `clear mex`: run the mexAtExit handlers (the gateways release their device context there)
int mm_clear() { int k = (int)g_atexit.size(); for (auto fn : g_atexit) fn(); g_atexit.clear(); return k; }
void mm_free_all() { for (auto* a : g_all) delete a; g_all.clear(); }
}
