/* oracle/ref_exit_hook.c -- TEST INFRASTRUCTURE ONLY.
 * A local definition of exit() for oracle/_ref/liblgar_ref.so (linked -Bsymbolic-functions), so
 * that the reference's exit() calls on numerical failure (reference src/lgar.cxx:2852) return
 * control to the harness instead of ending the pytest process. */
void lgref_exit_hook(int code);
void exit(int code) { lgref_exit_hook(code); for (;;) {} }
