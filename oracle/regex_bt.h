/* regex_bt.h — backtracking regex engine of the oracle (TEST INFRASTRUCTURE ONLY; see regex_bt.c). */
#ifndef ORX_REGEX_BT_H
#define ORX_REGEX_BT_H
#define _GNU_SOURCE
typedef struct orx_prog orx_prog;
orx_prog* orx_compile(const char* pattern);   /* NULL on unsupported syntax */
void orx_free(orx_prog* p);
int orx_num_caps(const orx_prog* p);
/* Unanchored leftmost-first search over s[0..n).  caps[2*i], caps[2*i+1] = byte span of group i (-1 if unset). */
int orx_search(const orx_prog* p, const char* s, int n, int* caps);
#endif
