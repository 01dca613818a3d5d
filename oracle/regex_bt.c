/* regex_bt.c — a small backtracking regular-expression engine for the oracle.
 *
 * TEST INFRASTRUCTURE ONLY (see oracle.c header).  The reference matches kmsg lines with Go's stdlib `regexp`
 * (RE2 syntax, leftmost-first semantics; call sites components/accelerator/nvidia/xid/kmsg.go:47-50,74,117,248,252 and
 * sxid/kmsg.go:24-25,32,43).  Go's regexp is not in /root/reference (toolchain `go 1.25.7`, absent here), so this file
 * restates the published algorithm: compile to a Thompson-style program (char / any / class / split / jmp / save /
 * match), run a priority-ordered depth-first search with a visited (pc, pos) bitmap — the same scheme as Go's
 * regexp/backtrack.go — which yields exactly the leftmost-first submatch.  The engine runs the reference's regex strings
 * VERBATIM; only the subset of RE2 syntax those six patterns use is supported (literals, escapes \( \) \. \d \s, '.',
 * bracket classes with ranges and negation, (...) and (?:...), '|', greedy and lazy * + ?, {n}, a leading (?s)).
 * Bytes are matched as bytes: every literal in the patterns is ASCII, so this equals RE2's UTF-8 behaviour.
 */
#include "regex_bt.h"

#include <stdlib.h>
#include <string.h>

enum { OP_CHAR, OP_ANY, OP_ANYNL, OP_CLASS, OP_SPLIT, OP_JMP, OP_SAVE, OP_MATCH };

typedef struct { int op, x, y; unsigned char c; } Inst;

struct orx_prog {
  Inst* code;
  int n, cap;
  unsigned char (*classes)[32];   /* 256-bit sets */
  int n_classes;
  int n_caps;                     /* capture groups incl. group 0 */
  int dotall;
  char prefix[64];                /* required literal prefix (Go's regexp does the same prefix acceleration) */
  int prefix_len;
};

/* ---------------- parser -> AST ---------------- */
typedef struct Node Node;
struct Node {
  int kind;            /* 'c' char, '.' any, '[' class, '(' group, '|' alt, '&' concat, 'q' quant, 'e' empty */
  unsigned char ch;
  int cls;             /* class index */
  int cap;             /* capture index or -1 */
  int min, max, lazy;  /* quant: max -1 = inf */
  Node *a, *b;
};

typedef struct { const char* s; int i, n; orx_prog* p; int err; } Parser;

static Node* mk(int kind) { Node* n = (Node*)calloc(1, sizeof(Node)); n->kind = kind; n->cap = -1; return n; }
static Node* parse_alt(Parser* P);

static void cls_set(unsigned char* set, int c) { set[c >> 3] |= (unsigned char)(1u << (c & 7)); }
static void cls_range(unsigned char* set, int lo, int hi) { for (int c = lo; c <= hi; ++c) cls_set(set, c); }
static void cls_escape(unsigned char* set, int e) {
  if (e == 'd') cls_range(set, '0', '9');
  else if (e == 's') { cls_set(set, '\t'); cls_set(set, '\n'); cls_set(set, '\f'); cls_set(set, '\r'); cls_set(set, ' '); }  /* RE2 \s */
  else cls_set(set, e);
}
static int new_class(orx_prog* p) {
  p->classes = realloc(p->classes, (size_t)(p->n_classes + 1) * 32);
  memset(p->classes[p->n_classes], 0, 32);
  return p->n_classes++;
}

static Node* parse_atom(Parser* P) {
  if (P->i >= P->n) return mk('e');
  char c = P->s[P->i];
  if (c == '(') {
    ++P->i;
    Node* g = mk('(');
    if (P->i + 1 < P->n && P->s[P->i] == '?' && P->s[P->i + 1] == ':') P->i += 2;
    else g->cap = P->p->n_caps++;
    g->a = parse_alt(P);
    if (P->i >= P->n || P->s[P->i] != ')') P->err = 1; else ++P->i;
    return g;
  }
  if (c == '[') {
    ++P->i;
    Node* k = mk('[');
    k->cls = new_class(P->p);
    unsigned char* set = P->p->classes[k->cls];
    int neg = 0;
    if (P->i < P->n && P->s[P->i] == '^') { neg = 1; ++P->i; }
    int first = 1;
    while (P->i < P->n && (P->s[P->i] != ']' || first)) {
      first = 0;
      int lo = (unsigned char)P->s[P->i++];
      if (lo == '\\' && P->i < P->n) {
        int e = (unsigned char)P->s[P->i++];
        if (e == 'd' || e == 's') { cls_escape(set, e); continue; }
        lo = e;
      }
      if (P->i + 1 < P->n && P->s[P->i] == '-' && P->s[P->i + 1] != ']') {
        int hi = (unsigned char)P->s[P->i + 1];
        P->i += 2;
        if (hi == '\\' && P->i < P->n) hi = (unsigned char)P->s[P->i++];
        cls_range(set, lo, hi);
      } else cls_set(set, lo);
    }
    if (P->i >= P->n) P->err = 1; else ++P->i;
    if (neg) for (int b = 0; b < 32; ++b) set[b] = (unsigned char)~set[b];
    return k;
  }
  if (c == '.') { ++P->i; return mk('.'); }
  if (c == '\\') {
    ++P->i;
    if (P->i >= P->n) { P->err = 1; return mk('e'); }
    int e = (unsigned char)P->s[P->i++];
    if (e == 'd' || e == 's' || e == 'S') {
      Node* k = mk('[');
      k->cls = new_class(P->p);
      cls_escape(P->p->classes[k->cls], e == 'S' ? 's' : e);
      if (e == 'S') for (int b = 0; b < 32; ++b) P->p->classes[k->cls][b] = (unsigned char)~P->p->classes[k->cls][b];   /* \S = [^\t\n\f\r ] */
      return k;
    }
    Node* l = mk('c');
    l->ch = (unsigned char)e;
    return l;
  }
  ++P->i;
  Node* l = mk('c');
  l->ch = (unsigned char)c;
  return l;
}

static Node* parse_repeat(Parser* P) {
  Node* a = parse_atom(P);
  while (P->i < P->n) {
    char c = P->s[P->i];
    int mn, mx;
    if (c == '*') { mn = 0; mx = -1; ++P->i; }
    else if (c == '+') { mn = 1; mx = -1; ++P->i; }
    else if (c == '?') { mn = 0; mx = 1; ++P->i; }
    else if (c == '{') {
      int j = P->i + 1, v = 0;
      while (j < P->n && P->s[j] >= '0' && P->s[j] <= '9') v = v * 10 + (P->s[j++] - '0');
      if (j >= P->n || P->s[j] != '}') { P->err = 1; return a; }
      P->i = j + 1;
      mn = mx = v;
    } else break;
    Node* q = mk('q');
    q->a = a; q->min = mn; q->max = mx;
    if (P->i < P->n && P->s[P->i] == '?') { q->lazy = 1; ++P->i; }
    a = q;
  }
  return a;
}

static Node* parse_concat(Parser* P) {
  Node* left = NULL;
  while (P->i < P->n && P->s[P->i] != '|' && P->s[P->i] != ')') {
    Node* r = parse_repeat(P);
    if (!left) left = r;
    else { Node* c = mk('&'); c->a = left; c->b = r; left = c; }
  }
  return left ? left : mk('e');
}

static Node* parse_alt(Parser* P) {
  Node* left = parse_concat(P);
  while (P->i < P->n && P->s[P->i] == '|') {
    ++P->i;
    Node* r = parse_concat(P);
    Node* a = mk('|');
    a->a = left; a->b = r;
    left = a;
  }
  return left;
}

/* ---------------- emitter ---------------- */
static int emit(orx_prog* p, int op, int x, int y, unsigned char c) {
  if (p->n == p->cap) { p->cap = p->cap ? p->cap * 2 : 64; p->code = realloc(p->code, (size_t)p->cap * sizeof(Inst)); }
  p->code[p->n] = (Inst){op, x, y, c};
  return p->n++;
}

static void gen(orx_prog* p, const Node* n) {
  switch (n->kind) {
    case 'e': break;
    case 'c': emit(p, OP_CHAR, 0, 0, n->ch); break;
    case '.': emit(p, p->dotall ? OP_ANYNL : OP_ANY, 0, 0, 0); break;
    case '[': emit(p, OP_CLASS, n->cls, 0, 0); break;
    case '&': gen(p, n->a); gen(p, n->b); break;
    case '(':
      if (n->cap >= 0) emit(p, OP_SAVE, 2 * n->cap, 0, 0);
      gen(p, n->a);
      if (n->cap >= 0) emit(p, OP_SAVE, 2 * n->cap + 1, 0, 0);
      break;
    case '|': {
      int s = emit(p, OP_SPLIT, 0, 0, 0);
      p->code[s].x = p->n;
      gen(p, n->a);
      int j = emit(p, OP_JMP, 0, 0, 0);
      p->code[s].y = p->n;
      gen(p, n->b);
      p->code[j].x = p->n;
      break;
    }
    case 'q': {
      for (int i = 0; i < n->min; ++i) gen(p, n->a);
      if (n->max < 0) {            /* e* : L1: split L2, L3; L2: e; jmp L1; L3:   (lazy swaps the preference) */
        int s = emit(p, OP_SPLIT, 0, 0, 0);
        int body = p->n;
        gen(p, n->a);
        emit(p, OP_JMP, s, 0, 0);
        if (n->lazy) { p->code[s].x = p->n; p->code[s].y = body; } else { p->code[s].x = body; p->code[s].y = p->n; }
      } else {
        for (int i = n->min; i < n->max; ++i) {   /* e? : split L1, L2; L1: e; L2: */
          int s = emit(p, OP_SPLIT, 0, 0, 0);
          int body = p->n;
          gen(p, n->a);
          if (n->lazy) { p->code[s].x = p->n; p->code[s].y = body; } else { p->code[s].x = body; p->code[s].y = p->n; }
        }
      }
      break;
    }
  }
}

static void free_node(Node* n) { if (!n) return; free_node(n->a); free_node(n->b); free(n); }

orx_prog* orx_compile(const char* pattern) {
  orx_prog* p = (orx_prog*)calloc(1, sizeof(orx_prog));
  Parser P = {pattern, 0, (int)strlen(pattern), p, 0};
  p->n_caps = 1;
  if (P.n >= 4 && !strncmp(pattern, "(?s)", 4)) { p->dotall = 1; P.i = 4; }
  Node* root = parse_alt(&P);
  if (P.err || P.i != P.n) { free_node(root); orx_free(p); return NULL; }
  emit(p, OP_SAVE, 0, 0, 0);
  gen(p, root);
  emit(p, OP_SAVE, 1, 0, 0);
  emit(p, OP_MATCH, 0, 0, 0);
  free_node(root);
  /* literal prefix = the leading run of CHAR instructions */
  for (int i = 1; i < p->n && p->code[i].op == OP_CHAR && p->prefix_len < 63; ++i) p->prefix[p->prefix_len++] = (char)p->code[i].c;
  return p;
}

void orx_free(orx_prog* p) {
  if (!p) return;
  free(p->code);
  free(p->classes);
  free(p);
}

int orx_num_caps(const orx_prog* p) { return p->n_caps; }

/* ---------------- matcher ---------------- */
typedef struct { int pc; int pos; int restore; } Job;

int orx_search(const orx_prog* p, const char* s, int n, int* caps /* 2 * n_caps */) {
  /* quick reject / candidate starts through the literal prefix, like Go's regexp prefix acceleration */
  const char* first = s;
  if (p->prefix_len) {
    first = (const char*)memmem(s, (size_t)n, p->prefix, (size_t)p->prefix_len);
    if (!first) return 0;
  }
  const size_t bits = (size_t)p->n * (size_t)(n + 1);
  unsigned char* visited = (unsigned char*)calloc((bits + 7) / 8, 1);
  int jcap = 256, jn = 0;
  Job* stack = (Job*)malloc((size_t)jcap * sizeof(Job));
  int found = 0;
  for (int i = 0; i < 2 * p->n_caps; ++i) caps[i] = -1;
  int start = (int)(first - s);
  while (start <= n && !found) {
    jn = 0;
    for (int i = 0; i < 2 * p->n_caps; ++i) caps[i] = -1;
    stack[jn++] = (Job){0, start, 0};
    while (jn && !found) {
      Job j = stack[--jn];
      if (j.restore) { caps[j.pc] = j.pos; continue; }
      int pc = j.pc, pos = j.pos;
      for (;;) {
        const size_t bit = (size_t)pc * (size_t)(n + 1) + (size_t)pos;
        if (visited[bit >> 3] & (1u << (bit & 7))) break;
        visited[bit >> 3] |= (unsigned char)(1u << (bit & 7));
        const Inst* in = &p->code[pc];
        if (in->op == OP_CHAR) { if (pos < n && (unsigned char)s[pos] == in->c) { ++pc; ++pos; continue; } break; }
        if (in->op == OP_ANY) { if (pos < n && s[pos] != '\n') { ++pc; ++pos; continue; } break; }
        if (in->op == OP_ANYNL) { if (pos < n) { ++pc; ++pos; continue; } break; }
        if (in->op == OP_CLASS) {
          if (pos < n && (p->classes[in->x][(unsigned char)s[pos] >> 3] >> ((unsigned char)s[pos] & 7)) & 1) { ++pc; ++pos; continue; }
          break;
        }
        if (in->op == OP_JMP) { pc = in->x; continue; }
        if (jn + 2 >= jcap) { jcap *= 2; stack = (Job*)realloc(stack, (size_t)jcap * sizeof(Job)); }
        if (in->op == OP_SPLIT) { stack[jn++] = (Job){in->y, pos, 0}; pc = in->x; continue; }
        if (in->op == OP_SAVE) { stack[jn++] = (Job){in->x, caps[in->x], 1}; caps[in->x] = pos; ++pc; continue; }
        if (in->op == OP_MATCH) { found = 1; break; }
      }
    }
    if (found) break;
    /* next candidate start */
    if (p->prefix_len) {
      if (start + 1 > n) break;
      const char* nx = (const char*)memmem(s + start + 1, (size_t)(n - start - 1), p->prefix, (size_t)p->prefix_len);
      if (!nx) break;
      start = (int)(nx - s);
    } else ++start;
  }
  free(stack);
  free(visited);
  return found;
}
