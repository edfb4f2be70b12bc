/* Hand-written stand-in for the autoconf-generated config.h of the reference
 * (x86-64 linux, gcc).  Test infrastructure only: used by oracle/ref/Makefile
 * to compile the reference sources in place into oracle/_ref/. */
#define HAVE_STDLIB_H 1
#define HAVE_UNISTD_H 1
#define HAVE_DLFCN_H 1
#define HAVE_TIMEGM 1
#define SIZEOF_SHORT 2
#define SIZEOF_INT 4
#define SIZEOF_LONG 8
#define SIZEOF_FLOAT 4
#define SIZEOF_DOUBLE 8
#define HAVE_ISNAN 1
#define HAVE_ISINF 1
#define HAVE_FINITE 1
#define HAVE_ERF 1
#define HAVE_FSEEKO 1
#define FSEEK(a,b,c) fseeko(a,b,c)
#define FTELL(a) ftello(a)
