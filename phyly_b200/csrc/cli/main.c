/*
 * arbplf-* executables: no arguments, one JSON document on stdin, one JSON
 * document on stdout, exit status = retcode (src/arbplf-ll.c:4-15 and
 * siblings; runjson.c:117-157).  Compiled once per program with
 * -DARBPLF_FN=<entry point>.
 */
#include "arbplf.h"

int main(void)
{
    return arbplf_run_stdio(ARBPLF_FN);
}
