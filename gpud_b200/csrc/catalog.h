// catalog.h — Xid / SXid / NVLink-rule lookup tables in the POD form the device kernels read.
//
// Built once per process on the host from the generated data (catalog_data.inc, derived mechanically from
// components/accelerator/nvidia/xid/xid.go:122-2952, xid/catalog_generated.go:7-277, sxid/sxid.go:94-2380)
// following the construction in xid/xid.go:2997-3090 (buildNVLinkSubCodeDetails + applyOperationalOverrides),
// then uploaded to each scanning device.
#pragma once
#include <stdint.h>

typedef struct { int code; int event; int n_actions; int actions[4]; const char* description; const char* mnemonic; } gpud_cat_xid_row;
typedef struct { int xid; const char* unit; const char* pat_v1; const char* pat_v2; uint32_t error_status; const char* resolution;
                 const char* investigatory; const char* severity; } gpud_cat_rule_row;
typedef struct { int sxid; int event; int n_actions; int actions[4]; int potential_fatal; int always_fatal; const char* name; } gpud_cat_sxid_row;

#define GPUD_T_MAX_XID 192
#define GPUD_T_MAX_RULES 128
#define GPUD_T_MAX_SXID 128
#define GPUD_T_MAX_SUB 128
#define GPUD_T_ALIAS_MAX 4
#define GPUD_T_ALIAS_LEN 48

typedef struct {
  int8_t present, event, n_actions /* -1 = nil */, pad;
  int8_t actions[4];
} gpud_t_detail;

typedef struct {
  int32_t xid;
  uint32_t error_status;
  uint32_t v1_care, v1_val, v2_care, v2_val;
  uint8_t v1_kind, v2_kind;          /* 0 empty (matches anything) / 1 valid 32-char pattern / 2 malformed (never matches) */
  int8_t rule_event;                 /* event from Severity, else from the Resolution bucket; 0 = Unknown               */
  int8_t rule_n_actions;             /* -1 = bucket maps to no action                                                   */
  int8_t rule_action;
  uint8_t has_hint;                  /* Investigatory is neither "", IGNORE nor CONTACT_SUPPORT                         */
  uint8_t n_alias;
  uint8_t pad;
  char alias[GPUD_T_ALIAS_MAX][GPUD_T_ALIAS_LEN]; /* normalized (upper, '-'->'_', [A-Z0-9_] only), NUL-terminated       */
} gpud_t_rule;

typedef struct {
  int32_t xid, sub_code;
  uint32_t error_status;   /* only meaningful for by_status rows */
  gpud_t_detail d;
  int32_t variant;         /* 0, or 1 / 2 for the 149.4 / 149.10 operational overrides */
} gpud_t_sub;

typedef struct { int32_t code; gpud_t_detail d; } gpud_t_sxid;

typedef struct {
  gpud_t_detail xid[GPUD_T_MAX_XID];        /* indexed by code */
  int32_t n_rules;
  gpud_t_rule rules[GPUD_T_MAX_RULES];
  int32_t n_by_status;
  gpud_t_sub by_status[GPUD_T_MAX_SUB];     /* detailsWithSubCodesByStatus */
  int32_t n_by_sub;
  gpud_t_sub by_sub[GPUD_T_MAX_SUB];        /* detailsWithSubCodes */
  uint8_t has_sub_map[GPUD_T_MAX_XID];      /* detailsWithSubCodes has an entry for this xid */
  int32_t n_sxid;
  gpud_t_sxid sxid[GPUD_T_MAX_SXID];        /* sorted by code */
} gpud_tables;

const gpud_tables* gpud_host_tables(void);   /* built on first use (thread-safe) */
