/*
 * pgshim/postgres.h -- TEST INFRASTRUCTURE ONLY.
 *
 * Minimal stand-in for the PostgreSQL server headers so that the reference
 * extension's kmer.c/kmer.h (read in place from /root/reference, never copied)
 * compile unmodified into oracle/_ref/libkmer_ref.so.  Only the symbols that
 * kmer.c actually uses are provided (list: SURVEY.md section 8c).
 *
 * Semantics kept from PostgreSQL (little-endian varlena):
 *   4-byte header  : uint32 = total_len << 2          (low two bits 00)
 *   1-byte header  : uint8  = (total_len << 1) | 1    (low bit 1)
 *   ereport(ERROR) : longjmp to the innermost handler installed by the driver,
 *                    with sqlstate / message / detail captured (thread-local).
 *   palloc         : bump arena, thread-local, reset by the driver per row.
 */
#ifndef PGSHIM_POSTGRES_H
#define PGSHIM_POSTGRES_H

#include <stdint.h>
#include <stdbool.h>
#include <stddef.h>
#include <stdlib.h>
#include <string.h>
#include <stdio.h>
#include <stdarg.h>
#include <setjmp.h>

typedef uintptr_t Datum;
typedef int16_t int16;
typedef int32_t int32;
typedef uint32_t uint32;
typedef uint8_t uint8;
typedef size_t Size;

struct varlena
{
	char vl_len_[4];
	char vl_dat[];
};

#define VARHDRSZ ((int32) sizeof(int32))
#define VARHDRSZ_SHORT 1

#define VARATT_IS_SHORT(p) ((((const uint8 *) (p))[0] & 0x01) == 0x01)
#define VARSIZE_4B(p) ((*(const uint32 *) (p)) >> 2)
#define VARSIZE_SHORT(p) ((((const uint8 *) (p))[0] >> 1) & 0x7F)
#define VARSIZE_ANY_EXHDR(p) \
	(VARATT_IS_SHORT(p) ? (int) VARSIZE_SHORT(p) - VARHDRSZ_SHORT : (int) VARSIZE_4B(p) - VARHDRSZ)
#define VARDATA_ANY(p) \
	(VARATT_IS_SHORT(p) ? ((char *) (p)) + VARHDRSZ_SHORT : ((char *) (p)) + VARHDRSZ)
#define SET_VARSIZE(p, len) (*(uint32 *) (p) = ((uint32) (len)) << 2)
#define SET_VARSIZE_SHORT(p, len) (((uint8 *) (p))[0] = (uint8) ((((uint32) (len)) << 1) | 0x01))

/* ---- memory ---- */
typedef struct PgShimArena *MemoryContext;
extern void *pgshim_palloc(Size n);
extern char *pgshim_pstrdup(const char *s);
extern char *pgshim_psprintf(const char *fmt, ...);
#define palloc(n) pgshim_palloc(n)
#define pstrdup(s) pgshim_pstrdup(s)
#define psprintf(...) pgshim_psprintf(__VA_ARGS__)
static inline MemoryContext MemoryContextSwitchTo(MemoryContext c) { return c; }

#define PointerGetDatum(p) ((Datum) (uintptr_t) (p))
#define DatumGetPointer(d) ((void *) (uintptr_t) (d))
#define Int64GetDatum(x) ((Datum) (int64_t) (x))
#define DatumGetInt64(d) ((int64_t) (d))
#define BoolGetDatum(x) ((Datum) ((x) ? 1 : 0))
#define DatumGetBool(d) ((bool) ((d) != 0))
#define PG_DETOAST_DATUM(d) ((struct varlena *) DatumGetPointer(d)) /* nothing is ever toasted here */
typedef int64_t int64;
extern void pgshim_pfree(void *p);
#define pfree(p) pgshim_pfree(p)
/* PG_TRY / PG_CATCH / PG_END_TRY / PG_RE_THROW over the same longjmp handler chain ereport() uses */
#define PG_TRY() do { jmp_buf *pgshim_saved_ = pgshim_handler; jmp_buf pgshim_local_; pgshim_handler = &pgshim_local_; if (setjmp(pgshim_local_) == 0) {
#define PG_CATCH() pgshim_handler = pgshim_saved_; } else { pgshim_handler = pgshim_saved_;
#define PG_END_TRY() } } while (0)
#define PG_RE_THROW() pgshim_throw()

/* ---- errors ---- */
#define ERROR 21
#define ERRCODE_INVALID_TEXT_REPRESENTATION 0x22503 /* "22P02" tag, value only compared for identity */
#define ERRCODE_STRING_DATA_RIGHT_TRUNCATION 0x22001
#define ERRCODE_INVALID_PARAMETER_VALUE 0x22023
#define ERRCODE_INTERNAL_ERROR 0x99000 /* XX000 */
#define ERRCODE_OUT_OF_MEMORY 0x53200

typedef struct PgShimError
{
	int sqlstate;
	char message[128];
	char detail[128];
} PgShimError;

extern __thread PgShimError pgshim_error;
extern __thread jmp_buf *pgshim_handler;

extern int pgshim_errcode(int code);
extern int pgshim_errmsg(const char *fmt, ...);
extern int pgshim_errdetail(const char *fmt, ...);
extern void pgshim_throw(void) __attribute__((noreturn));

#define errcode(c) pgshim_errcode(c)
#define errmsg(...) pgshim_errmsg(__VA_ARGS__)
#define errdetail(...) pgshim_errdetail(__VA_ARGS__)
/* ereport(ERROR, (errcode(..), errmsg(..), ...)) : evaluate the aux calls, then unwind */
#define ereport(elevel, rest) \
	do { pgshim_error.sqlstate = 0; pgshim_error.message[0] = 0; pgshim_error.detail[0] = 0; \
	     (void) (rest); if ((elevel) >= ERROR) pgshim_throw(); } while (0)

#endif
