// catalog.cpp — host-side construction of the device lookup tables and the catalog string accessors.
// Follows xid/xid.go:2997-3090 (sub-code detail maps + operational overrides), :3099-3218 (rule matching inputs),
// :3222-3283 (severity / bucket -> event type, bucket -> action).
#include "catalog.h"

#include <ctype.h>
#include <stdio.h>
#include <string.h>

#include <algorithm>
#include <mutex>
#include <string>
#include <vector>

#include "../../include/gpud_b200.h"
#include "catalog_data.inc"

namespace {

int event_from_severity(const char* s) {        // xid.go:3263-3272
  std::string v(s);
  while (!v.empty() && isspace((unsigned char)v.back())) v.pop_back();
  size_t b = 0;
  while (b < v.size() && isspace((unsigned char)v[b])) ++b;
  v = v.substr(b);
  for (auto& c : v) c = (char)tolower((unsigned char)c);
  if (v == "fatal" || v == "fatal**" || v == "link fatal" || v == "link fatal?") return GPUD_EVENT_FATAL;
  if (v == "non-fatal" || v == "non-fatal*") return GPUD_EVENT_WARNING;
  return GPUD_EVENT_UNKNOWN;
}

bool in_list(const char* s, std::initializer_list<const char*> l) {
  for (const char* x : l)
    if (!strcmp(s, x)) return true;
  return false;
}

int event_from_bucket(const char* b) {          // xid.go:3222-3246
  if (in_list(b, {"CONTACT_SUPPORT", "CHECK_MECHANICALS", "WORKFLOW_NVLINK_ERR", "WORKFLOW_NVLINK5_ERR", "XID_154", "XID_154_EVAL", "RESTART_BM"}))
    return GPUD_EVENT_FATAL;
  if (in_list(b, {"RESET_GPU", "RESTART_APP", "RESTART_VM", "CHECK_UVM", "WORKFLOW_XID_48", "WORKFLOW_XID_45", "UPDATE_SWFW"}))
    return GPUD_EVENT_CRITICAL;
  if (in_list(b, {"IGNORE", ""})) return GPUD_EVENT_INFO;
  return GPUD_EVENT_WARNING;
}

int action_from_bucket(const char* b) {         // xid.go:3248-3261 ; 0 = none
  if (in_list(b, {"CONTACT_SUPPORT", "CHECK_MECHANICALS", "WORKFLOW_NVLINK_ERR", "WORKFLOW_NVLINK5_ERR", "XID_154", "XID_154_EVAL"}))
    return GPUD_ACT_HARDWARE_INSPECTION;
  if (in_list(b, {"RESET_GPU", "RESTART_BM", "RESTART_VM", "CHECK_UVM"})) return GPUD_ACT_REBOOT_SYSTEM;
  if (in_list(b, {"RESTART_APP", "WORKFLOW_XID_45", "WORKFLOW_XID_48", "UPDATE_SWFW"})) return GPUD_ACT_CHECK_USER_APP_AND_GPU;
  if (in_list(b, {"IGNORE", ""})) return GPUD_ACT_IGNORE_NO_ACTION_REQUIRED;
  return 0;
}

const char* action_wire(int a) {                // api/v1/types.go:183-203
  switch (a) {
    case GPUD_ACT_IGNORE_NO_ACTION_REQUIRED: return "IGNORE_NO_ACTION_REQUIRED";
    case GPUD_ACT_REBOOT_SYSTEM: return "REBOOT_SYSTEM";
    case GPUD_ACT_HARDWARE_INSPECTION: return "HARDWARE_INSPECTION";
    case GPUD_ACT_CHECK_USER_APP_AND_GPU: return "CHECK_USER_APP_AND_GPU";
  }
  return "";
}

// pattern "0/1/-" x32, MSB first -> (care, value); kind per catalog.h
void compile_pattern(const char* p, uint32_t* care, uint32_t* val, uint8_t* kind) {
  *care = *val = 0;
  const size_t n = strlen(p);
  if (n == 0) { *kind = 0; return; }
  if (n != 32) { *kind = 2; return; }
  for (int i = 0; i < 32; ++i) {
    const int bit = 31 - i;
    if (p[i] == '1') { *care |= 1u << bit; *val |= 1u << bit; }
    else if (p[i] == '0') { *care |= 1u << bit; }
    else if (p[i] != '-') { *kind = 2; return; }
  }
  *kind = 1;
}

// sampleFromPattern (xid.go:3127-3145): value with '1' bits set, false if empty / malformed
bool sample_from_pattern(const char* p, uint32_t* v) {
  uint32_t care, val;
  uint8_t kind;
  compile_pattern(p, &care, &val, &kind);
  if (kind != 1) return false;
  *v = val;
  return true;
}

std::string normalize_unit(const std::string& s) {   // xid.go:3203-3218
  size_t b = 0, e = s.size();
  while (b < e && isspace((unsigned char)s[b])) ++b;
  while (e > b && isspace((unsigned char)s[e - 1])) --e;
  std::string o;
  for (size_t i = b; i < e; ++i) {
    char c = (char)toupper((unsigned char)s[i]);
    if (c == '-') c = '_';
    if ((c >= 'A' && c <= 'Z') || (c >= '0' && c <= '9') || c == '_') o.push_back(c);
  }
  return o;
}

std::vector<std::string> unit_aliases(const std::string& u) {   // xid.go:3188-3201
  std::vector<std::string> al;
  std::string cur;
  for (char c : u) {
    if (c == '/' || c == ',' || c == '(' || c == ')' || c == ' ') {
      if (!cur.empty()) al.push_back(cur);
      cur.clear();
    } else cur.push_back(c);
  }
  if (!cur.empty()) al.push_back(cur);
  if (al.empty()) al.push_back(u);
  al.push_back(u);
  return al;
}

gpud_t_detail make_detail(int event, int n_actions, const int* acts) {
  gpud_t_detail d;
  memset(&d, 0, sizeof d);
  d.present = 1;
  d.event = (int8_t)event;
  d.n_actions = (int8_t)n_actions;
  for (int i = 0; i < 4 && i < n_actions; ++i) d.actions[i] = (int8_t)acts[i];
  return d;
}

// mergeSuggestedActions (xid.go:3317-3336): set union, sorted by the wire string
void merge_actions(gpud_t_detail* base, const gpud_t_detail& add) {
  if (base->n_actions < 0) { base->n_actions = add.n_actions; memcpy(base->actions, add.actions, 4); return; }
  if (add.n_actions < 0) return;
  std::vector<int> set;
  for (int i = 0; i < base->n_actions; ++i) set.push_back(base->actions[i]);
  for (int i = 0; i < add.n_actions; ++i) set.push_back(add.actions[i]);
  std::sort(set.begin(), set.end(), [](int a, int b) { return strcmp(action_wire(a), action_wire(b)) < 0; });
  set.erase(std::unique(set.begin(), set.end()), set.end());
  base->n_actions = (int8_t)std::min<size_t>(4, set.size());
  for (int i = 0; i < base->n_actions; ++i) base->actions[i] = (int8_t)set[i];
}

gpud_tables g_tables;
std::once_flag g_once;

gpud_t_sub* find_sub(gpud_t_sub* arr, int n, int xid, int sc, bool with_status, uint32_t st) {
  for (int i = 0; i < n; ++i)
    if (arr[i].xid == xid && arr[i].sub_code == sc && (!with_status || arr[i].error_status == st)) return &arr[i];
  return nullptr;
}

void build_tables() {
  gpud_tables& T = g_tables;
  memset(&T, 0, sizeof T);
  for (int i = 0; i < GPUD_CAT_N_XID; ++i) {
    const gpud_cat_xid_row& r = GPUD_CAT_XID[i];
    // SuggestedActionsByGPUd == nil is encoded as n_actions -1
    T.xid[r.code] = make_detail(r.event, r.n_actions > 0 ? r.n_actions : -1, r.actions);
  }
  T.n_rules = GPUD_CAT_N_RULES;
  for (int i = 0; i < GPUD_CAT_N_RULES; ++i) {
    const gpud_cat_rule_row& r = GPUD_CAT_RULES[i];
    gpud_t_rule& o = T.rules[i];
    o.xid = r.xid;
    o.error_status = r.error_status;
    compile_pattern(r.pat_v1, &o.v1_care, &o.v1_val, &o.v1_kind);
    compile_pattern(r.pat_v2, &o.v2_care, &o.v2_val, &o.v2_kind);
    int ev = event_from_severity(r.severity);
    if (ev == GPUD_EVENT_UNKNOWN) ev = event_from_bucket(r.resolution);
    o.rule_event = (int8_t)ev;
    const int act = action_from_bucket(r.resolution);
    o.rule_n_actions = act ? 1 : -1;
    o.rule_action = (int8_t)act;
    o.has_hint = (r.investigatory[0] && strcmp(r.investigatory, "IGNORE") && strcmp(r.investigatory, "CONTACT_SUPPORT")) ? 1 : 0;
    std::vector<std::string> al;
    for (auto& a : unit_aliases(r.unit)) {
      std::string n = normalize_unit(a);
      if (std::find(al.begin(), al.end(), n) == al.end()) al.push_back(n);
    }
    o.n_alias = 0;
    for (auto& a : al) {
      if (o.n_alias >= GPUD_T_ALIAS_MAX || a.size() >= GPUD_T_ALIAS_LEN) continue;   // sizes asserted by the generator's data
      snprintf(o.alias[o.n_alias++], GPUD_T_ALIAS_LEN, "%s", a.c_str());
    }
  }
  // buildNVLinkSubCodeDetails (xid.go:2997-3060)
  for (int i = 0; i < GPUD_CAT_N_RULES; ++i) {
    const gpud_cat_rule_row& r = GPUD_CAT_RULES[i];
    if (r.xid < 144 || r.xid > 150) continue;
    uint32_t sample;
    int sc;
    if (sample_from_pattern(r.pat_v2, &sample)) sc = (int)((sample >> 20) & 0x3F);
    else if (sample_from_pattern(r.pat_v1, &sample)) sc = (int)((sample >> 20) & 0x3F);
    else continue;
    T.has_sub_map[r.xid] = 1;
    const gpud_t_detail base = T.xid[r.xid];
    if (!base.present) continue;
    gpud_t_detail d = base;
    if (T.rules[i].rule_event != GPUD_EVENT_UNKNOWN) d.event = T.rules[i].rule_event;
    if (T.rules[i].rule_n_actions > 0) { d.n_actions = 1; memset(d.actions, 0, 4); d.actions[0] = T.rules[i].rule_action; }
    gpud_t_sub* ex = find_sub(T.by_status, T.n_by_status, r.xid, sc, true, r.error_status);
    if (ex) {
      d.event = std::max(ex->d.event, d.event);                    // maxEventType: rank == numeric id
      gpud_t_detail merged = ex->d;
      merge_actions(&merged, d);
      d.n_actions = merged.n_actions;
      memcpy(d.actions, merged.actions, 4);
      ex->d = d;
    } else {
      gpud_t_sub& s = T.by_status[T.n_by_status++];
      s.xid = r.xid; s.sub_code = sc; s.error_status = r.error_status; s.d = d; s.variant = 0;
    }
    gpud_t_sub* agg = find_sub(T.by_sub, T.n_by_sub, r.xid, sc, false, 0);
    if (!agg) {
      agg = &T.by_sub[T.n_by_sub++];
      agg->xid = r.xid; agg->sub_code = sc; agg->error_status = 0; agg->d = base; agg->variant = 0;
    }
    merge_actions(&agg->d, d);
  }
  // applyOperationalOverrides (xid.go:3062-3089)
  const int over[2][2] = {{4, 1}, {10, 2}};
  for (auto& ov : over) {
    gpud_t_sub* s = find_sub(T.by_sub, T.n_by_sub, 149, ov[0], false, 0);
    if (!s) continue;
    const int hw = GPUD_ACT_HARDWARE_INSPECTION;
    s->d = make_detail(GPUD_EVENT_FATAL, 1, &hw);
    s->variant = ov[1];
    for (int i = 0; i < T.n_by_status; ++i)
      if (T.by_status[i].xid == 149 && T.by_status[i].sub_code == ov[0]) { T.by_status[i].d = s->d; T.by_status[i].variant = ov[1]; }
  }
  T.n_sxid = GPUD_CAT_N_SXID;
  for (int i = 0; i < GPUD_CAT_N_SXID; ++i) {
    const gpud_cat_sxid_row& r = GPUD_CAT_SXID[i];
    T.sxid[i].code = r.sxid;
    T.sxid[i].d = make_detail(r.event, r.n_actions > 0 ? r.n_actions : -1, r.actions);
  }
  std::sort(T.sxid, T.sxid + T.n_sxid, [](const gpud_t_sxid& a, const gpud_t_sxid& b) { return a.code < b.code; });
}

const gpud_cat_xid_row* xid_row(int code) {
  for (int i = 0; i < GPUD_CAT_N_XID; ++i)
    if (GPUD_CAT_XID[i].code == code) return &GPUD_CAT_XID[i];
  return nullptr;
}

void json_escape(std::string& o, const char* s) {
  for (; *s; ++s) {
    unsigned char c = (unsigned char)*s;
    switch (c) {   // encoding/json escapes: \" \\ \n \r \t, <,>,& as \u00XX, control chars as \u00XX
      case '"': o += "\\\""; break;
      case '\\': o += "\\\\"; break;
      case '\n': o += "\\n"; break;
      case '\r': o += "\\r"; break;
      case '\t': o += "\\t"; break;
      case '<': case '>': case '&': { char b[8]; snprintf(b, sizeof b, "\\u%04x", c); o += b; break; }
      default:
        if (c < 0x20) { char b[8]; snprintf(b, sizeof b, "\\u%04x", c); o += b; }
        else o.push_back((char)c);
    }
  }
}

}  // namespace

const gpud_tables* gpud_host_tables(void) {
  std::call_once(g_once, build_tables);
  return &g_tables;
}

extern "C" const char* gpud_xid_description(int32_t code, int32_t variant) {
  if (variant == 1) return "NVLINK: NETIR Link Event - Possible NVLink cartridge error (contact provider)";        // xid.go:3067
  if (variant == 2) return "NVLINK: NETIR Link Event - Physical layer retransmission timeout (contact provider)";  // xid.go:3080
  const gpud_cat_xid_row* r = xid_row(code);
  return r ? r->description : "";
}
extern "C" const char* gpud_xid_mnemonic(int32_t code) {
  const gpud_cat_xid_row* r = xid_row(code);
  return r ? r->mnemonic : "";
}
extern "C" const char* gpud_sxid_name(int32_t code) {
  for (int i = 0; i < GPUD_CAT_N_SXID; ++i)
    if (GPUD_CAT_SXID[i].sxid == code) return GPUD_CAT_SXID[i].name;
  return "";
}
extern "C" const char* gpud_nvlink_rule_hint(int32_t idx) {
  if (idx < 0 || idx >= GPUD_CAT_N_RULES) return "";
  const char* h = GPUD_CAT_RULES[idx].investigatory;
  if (!h[0] || !strcmp(h, "IGNORE") || !strcmp(h, "CONTACT_SUPPORT")) return "";
  return h;
}

// (component, eventName, message) of the stateless kmsg line matchers, by hit kind.  nccl/kmsg_matcher.go:11-13,
// peermem/kmsg_matcher.go:13-15, infiniband/kmsg_matcher.go:14-58, cpu/kmsg_matcher.go:17-31, os/kmsg_matcher.go:17-19,
// disk/kmsg_matcher.go:10-56.
namespace {
struct KmsgEvent { const char *component, *event, *message; };
const KmsgEvent kKmsgEvents[GPUD_KIND_COUNT] = {
    {"", "", ""}, {"", "", ""}, {"", "", ""},
    {"nccl", "nvidia_nccl_segfault_in_libnccl", "NCCL communication error (segfault in libnccl.so)"},
    {"peermem", "nvidia_peermem_invalid_context", "peermem error detected (possible GPU communication issue)"},
    {"infiniband", "pci_power_insufficient", "Insufficient power on MLX5 PCIe slot"},
    {"infiniband", "port_module_high_temperature", "Overheated MLX5 adapter"},
    {"infiniband", "access_reg_failed", "MLX5 ACCESS_REG command failed - device may have restricted PF access"},
    {"cpu", "cpu_blocked_too_long", "CPU task blocked for more than 120 seconds"},
    {"cpu", "cpu_soft_lockup", "CPU soft lockup detected, not releasing for a period of time"},
    {"os", "vfs_file_max_limit_reached", "VFS file-max limit reached"},
    {"disk", "raid_array_failure", "RAID array has failed due to disk failure"},
    {"disk", "filesystem_read_only", "filesystem remounted as read-only due to errors"},
    {"disk", "nvme_path_failure", "NVMe device has no available path, I/O failing"},
    {"disk", "nvme_controller_timeout", "NVME controller I/O timeout detected, attempting reset"},
    {"disk", "nvme_device_disabled", "NVME device disabled after reset failure"},
    {"disk", "beyond_end_of_device", "I/O attempt beyond device boundaries detected"},
    {"disk", "buffer_io_error", "Buffer I/O error detected on device"},
    {"disk", "superblock_write_error", "I/O error while writing superblock"},
    // line primitives of the stateful matchers: no event of their own (gpud_kmsg_stateful_feed assembles "kernel_panic" / "OOM")
    {"os", "", ""}, {"os", "", ""}, {"memory", "", ""}, {"memory", "", ""}, {"memory", "", ""}, {"memory", "", ""},
};
const KmsgEvent& kmsg_event(int32_t kind) { return kKmsgEvents[(kind > 0 && kind < GPUD_KIND_COUNT) ? kind : 0]; }
}  // namespace
extern "C" const char* gpud_kmsg_event_name(int32_t kind) { return kmsg_event(kind).event; }
extern "C" const char* gpud_kmsg_event_message(int32_t kind) { return kmsg_event(kind).message; }
extern "C" const char* gpud_kmsg_component(int32_t kind) { return kmsg_event(kind).component; }
extern "C" int32_t gpud_kmsg_hit_message(const gpud_xid_hit* h, const uint8_t* buf, char* out, int32_t cap) {
  if (!h || !out || cap <= 0) return -1;
  std::string m = kmsg_event(h->kind).message;
  std::string capture;
  if (h->dev_len > 0) {
    if ((h->flags & GPUD_HIT_DEV_TRUNCATED) && buf) capture.assign((const char*)buf + h->dev_off, (size_t)h->dev_len);
    else capture.assign(h->device, strnlen(h->device, sizeof h->device));
  }
  if (h->kind == GPUD_KIND_CPU_BLOCKED_TOO_LONG || h->kind == GPUD_KIND_CPU_SOFT_LOCKUP) m += " (" + capture + ")";      // cpu/kmsg_matcher.go:54-63
  else if (h->kind == GPUD_KIND_IB_ACCESS_REG_FAILED && !capture.empty()) m += " (PCI device " + capture + ")";           // infiniband/kmsg_matcher.go:136-142
  if ((int32_t)m.size() + 1 > cap) return -1;
  memcpy(out, m.c_str(), m.size() + 1);
  return (int32_t)m.size();
}

// Reason of the sxid component's state (sxid/health_state.go:93-106): "SXID %d(%s) detected on %s" with the catalog name when the
// code is known, "SXID %d detected on %s" otherwise; sxid < 0 = no error -> "SXIDComponent is healthy".
extern "C" int32_t gpud_sxid_reason(int64_t sxid, const char* device, char* out, int32_t cap) {
  if (!out || cap <= 0) return -1;
  int n;
  if (sxid < 0) n = snprintf(out, (size_t)cap, "SXIDComponent is healthy");
  else {
    const char* name = (sxid <= 0x7fffffff) ? gpud_sxid_name((int32_t)sxid) : "";
    if (name && *name) n = snprintf(out, (size_t)cap, "SXID %lld(%s) detected on %s", (long long)sxid, name, device ? device : "");
    else n = snprintf(out, (size_t)cap, "SXID %lld detected on %s", (long long)sxid, device ? device : "");
  }
  return n < cap ? n : -1;
}

// collectFabricState's report over the box (fabric-manager/fabric_state.go:67-113): *healthy = no GPU has an issue; the reason is
// "GPU <uuid>: <issues>" per affected GPU, sorted, joined by "; ".
extern "C" int32_t gpud_fabric_report_reason(const gpud_fabric_raw* gpus, const char* const* gpu_uuids, int32_t n, int32_t* healthy, char* out, int32_t cap) {
  if (n < 0 || (n && (!gpus || !gpu_uuids)) || !out || cap <= 0) return -1;
  std::vector<std::string> reasons;
  char buf[512];
  for (int32_t i = 0; i < n; ++i) {
    const int32_t k = gpud_fabric_issues(&gpus[i], buf, sizeof buf);
    if (k < 0) return -1;
    if (k > 0) reasons.push_back(std::string("GPU ") + (gpu_uuids[i] ? gpu_uuids[i] : "") + ": " + buf);
  }
  std::sort(reasons.begin(), reasons.end());
  std::string j;
  for (size_t i = 0; i < reasons.size(); ++i) { if (i) j += "; "; j += reasons[i]; }
  if (healthy) *healthy = reasons.empty() ? 1 : 0;
  if ((int32_t)j.size() + 1 > cap) return -1;
  memcpy(out, j.c_str(), j.size() + 1);
  return (int32_t)j.size();
}

// GetDetail (xid/xid.go:74-77): the base catalog entry
extern "C" int32_t gpud_xid_get_detail(int32_t xid, int32_t* event_type, int32_t* n_actions, int32_t* actions4) {
  const gpud_tables* T = gpud_host_tables();
  if (xid < 0 || xid >= GPUD_T_MAX_XID || !T->xid[xid].present) return 0;
  const gpud_t_detail& d = T->xid[xid];
  if (event_type) *event_type = d.event;
  if (n_actions) *n_actions = d.n_actions;
  if (actions4) for (int i = 0; i < 4; ++i) actions4[i] = i < d.n_actions ? d.actions[i] : 0;
  return 1;
}

// sxid.GetDetail (sxid/sxid.go:31-35): 1 = known; n_actions -1 = SuggestedActionsByGPUd nil
extern "C" int32_t gpud_sxid_get_detail(int32_t sxid, int32_t* event_type, int32_t* n_actions, int32_t* actions4) {
  const gpud_tables* T = gpud_host_tables();
  for (int i = 0; i < T->n_sxid; ++i) {
    if (T->sxid[i].code != sxid) continue;
    const gpud_t_detail& d = T->sxid[i].d;
    if (event_type) *event_type = d.event;
    if (n_actions) *n_actions = d.n_actions;
    if (actions4) for (int k = 0; k < 4; ++k) actions4[k] = k < d.n_actions ? d.actions[k] : 0;
    return 1;
  }
  return 0;
}

// getDetailWithSubCodeAndStatus (xid/xid.go:97-107) with its fallbacks getDetailWithSubCode (:79-93) and GetDetail (:74-77), over
// the merged sub-code tables buildNVLinkSubCodeDetails made: 1 = found.  n_actions -1 = SuggestedActionsByGPUd nil;
// detail_variant selects the description (gpud_xid_description); sub_code_out = the Detail's SubCode.
extern "C" int32_t gpud_xid_detail(int32_t xid, int32_t sub_code, uint32_t error_status, int32_t* event_type, int32_t* n_actions, int32_t* actions4,
                                   int32_t* detail_variant, int32_t* sub_code_out) {
  const gpud_tables* T = gpud_host_tables();
  const gpud_t_detail* d = nullptr;
  int variant = 0, sc_out = 0;
  if (xid >= 0 && xid < GPUD_T_MAX_XID) {
    for (int i = 0; i < T->n_by_status && !d; ++i)
      if (T->by_status[i].xid == xid && T->by_status[i].sub_code == sub_code && T->by_status[i].error_status == error_status) {
        d = &T->by_status[i].d; variant = T->by_status[i].variant; sc_out = sub_code;
      }
    if (!d && T->has_sub_map[xid]) {
      for (int pass = 0; pass < 2 && !d; ++pass) {
        const int want = pass == 0 ? sub_code : 0;
        for (int i = 0; i < T->n_by_sub && !d; ++i)
          if (T->by_sub[i].xid == xid && T->by_sub[i].sub_code == want) { d = &T->by_sub[i].d; variant = T->by_sub[i].variant; sc_out = want; }
      }
    }
    if (!d && T->xid[xid].present) d = &T->xid[xid];
  }
  if (!d) return 0;
  if (event_type) *event_type = d->event;
  if (n_actions) *n_actions = d->n_actions;
  if (actions4) for (int i = 0; i < 4; ++i) actions4[i] = i < d->n_actions ? d->actions[i] : 0;
  if (detail_variant) *detail_variant = variant;
  if (sub_code_out) *sub_code_out = sc_out;
  return 1;
}

// (*xidErrorEventDetail).buildMessage (xid/health_state.go:130-169): the Message of a resolved xid event and the Reason of the
// component's health state.  NVLink codes 144-150 always carry the dotted sub-code and the error status; the text is the catalog
// mnemonic, followed by the description unless that is empty, "Unused" or the mnemonic itself; without a mnemonic the
// description alone.  gpu_uuid = what convertBusIDToUUID found ("" / NULL: none).  Returns the length, -1 if it does not fit.
extern "C" int32_t gpud_xid_build_message(uint64_t xid, int32_t sub_code, uint32_t error_status, const char* description, const char* device_uuid,
                                          const char* gpu_uuid, char* out, int32_t cap) {
  if (!out || cap <= 0) return -1;
  const char* dev = device_uuid ? device_uuid : "";
  const char* descr = description ? description : "";
  char header[96];
  if (xid >= 144 && xid <= 150) snprintf(header, sizeof header, "XID %llu.%d (err status 0x%08x)", (unsigned long long)xid, sub_code, error_status);
  else snprintf(header, sizeof header, "XID %llu", (unsigned long long)xid);
  std::string o = header;
  if (xid > (uint64_t)INT64_MAX) {                                   // intFromUint64 fails (xid/convert.go): no catalog lookup
    o += " detected on GPU ";
    o += dev;
  } else {
    std::string desc = xid <= 0x7fffffff ? gpud_xid_mnemonic((int32_t)xid) : "";
    if (desc.empty()) desc = descr;
    else if (descr[0] && strcmp(descr, "Unused") != 0 && desc != descr) { desc += " "; desc += descr; }
    o += " " + desc + " detected on GPU " + dev;
    if (gpu_uuid && *gpu_uuid) { o += " UUID:"; o += gpu_uuid; }
  }
  if ((int)o.size() + 1 > cap) return -1;
  memcpy(out, o.c_str(), o.size() + 1);
  return (int32_t)o.size();
}

// buildMessage for the payload the xid component persists for this hit (xid/component.go:503-521: sub-code, error status and
// description of Match's Detail).
extern "C" int32_t gpud_xid_hit_message(const gpud_xid_hit* h, const char* gpu_uuid, char* out, int32_t cap) {
  if (!h || h->kind != GPUD_KIND_XID) return -1;
  const bool ext = (h->flags & GPUD_HIT_EXTENDED) != 0;
  char dev[sizeof h->device + 1];
  memcpy(dev, h->device, sizeof h->device);
  dev[sizeof h->device] = 0;
  return gpud_xid_build_message((uint64_t)h->code, ext ? h->sub_code : 0, ext ? h->error_status : 0u, gpud_xid_description(h->code, h->detail_variant), dev,
                                gpu_uuid, out, cap);
}

// The test inside convertBusIDToUUID (xid/health_state.go:171-182): does the NVML device with this PCI bus id ("0000:04:00.0")
// belong to the device id of an xid event ("PCI:0000:04:00")?
extern "C" int32_t gpud_xid_device_matches_bus_id(const char* device_uuid, const char* pci_bus_id) {
  if (!device_uuid || !pci_bus_id) return 0;
  std::string want = device_uuid;
  if (want.compare(0, 4, "PCI:") == 0) want.erase(0, 4);
  want += ".";
  return strncmp(pci_bus_id, want.c_str(), want.size()) == 0 ? 1 : 0;
}

// ---- GPU product capabilities (pkg/nvidia/product/capabilities.go:6-137): decided from the NVML product name ----
namespace {
std::string lower(const char* s) {
  std::string o = s ? s : "";
  for (char& c : o) if (c >= 'A' && c <= 'Z') c = (char)(c - 'A' + 'a');
  return o;
}
struct ProductRow { const char* key; int mem_caps; int fm; };     // mem_caps: 1 containment | 2 page offlining | 4 row remapping; -1 = not in that map
const ProductRow kProducts[] = {                                   // gpuProductToMemMgmtCaps (:15-23), gpuProductToFMSupported (:25-50)
    {"a100", 7, 1}, {"b100", 7, 1}, {"b200", 7, 1}, {"gb200", 7, 0}, {"h100", 7, 1}, {"h200", 7, 1}, {"a10", 4, 0}, {"gh200", -1, 0},
};
// the longest key contained in the lowered name decides (:65-75, :121-134)
const ProductRow* longest_product(const std::string& p, bool mem_map) {
  const ProductRow* best = nullptr;
  for (const ProductRow& r : kProducts) {
    if (mem_map && r.mem_caps < 0) continue;
    if (p.find(r.key) == std::string::npos) continue;
    if (!best || strlen(best->key) < strlen(r.key)) best = &r;
  }
  return best;
}
}  // namespace

// SupportedMemoryMgmtCapsByGPUProduct (:119-137): 1 ErrorContainment | 2 DynamicPageOfflining | 4 RowRemapping
extern "C" int32_t gpud_product_mem_caps(const char* product_name) {
  const ProductRow* r = longest_product(lower(product_name), true);
  return r ? r->mem_caps : 0;
}
// SupportedFMByGPUProduct (:56-76): the on-node nv-fabricmanager daemon; PCIe variants never
extern "C" int32_t gpud_product_fm_supported(const char* product_name) {
  const std::string p = lower(product_name);
  if (p.find("pcie") != std::string::npos) return 0;
  const ProductRow* r = longest_product(p, false);
  return r ? r->fm : 0;
}
// SupportFabricStateByGPUProduct (:93-116): NVML fabric-state telemetry
extern "C" int32_t gpud_product_fabric_state_supported(const char* product_name) {
  const std::string p = lower(product_name);
  if (p.find("pcie") != std::string::npos || p.find("gh200") != std::string::npos) return 0;
  return (p.find("gb200") != std::string::npos || p.find("h100") != std::string::npos || p.find("h200") != std::string::npos) ? 1 : 0;
}

// The reason string of the nvlink component's check result for a verdict (nvlink/evaluate_threshold.go:11-35,77-188;
// component.go:307 for the no-issue case).  gpu_uuids[i] names gpu_index i (the reference lists UUIDs in sorted order, which is the
// gpu_index order); a missing name is rendered as "GPU-<index>".
extern "C" int32_t gpud_fabric_reason(const gpud_fabric_verdict* v, const char* const* gpu_uuids, int32_t n_uuids, char* out, int32_t cap) {
  if (!v || !out || cap <= 0 || n_uuids < 0) return -1;
  auto uuid = [&](int i) -> std::string {
    if (gpu_uuids && i < n_uuids && gpu_uuids[i] && *gpu_uuids[i]) return gpu_uuids[i];
    return "GPU-" + std::to_string(i);
  };
  auto list = [&](uint32_t mask) {
    std::string o;
    for (int i = 0; i < GPUD_MAX_GPUS; ++i)
      if (mask & (1u << i)) { if (!o.empty()) o += ","; o += uuid(i); }
    return o;
  };
  // appendNVLinkFailureDetails (:16-35)
  auto details = [&](std::string reason) {
    std::vector<std::string> parts;
    if (v->p2p_probed_pairs > 0 && v->p2p_ok_pairs == 0 && v->p2p_observed_status_mask != 0) {
      static const char* const kCode[7] = {"OK", "CNS", "GNS", "TNS", "DR", "NS", "U"};      // p2p.go:11-19
      std::vector<std::string> codes;
      for (int c = 0; c < 7; ++c) if (v->p2p_observed_status_mask & (1u << c)) codes.push_back(kCode[c]);
      std::sort(codes.begin(), codes.end());                                                // sortedKeys
      std::string j;
      for (size_t i = 0; i < codes.size(); ++i) { if (i) j += ","; j += codes[i]; }
      parts.push_back("peer nvlink p2p statuses=" + j);
    }
    if (v->inactive_mask) parts.push_back("inactive nvlinks=" + list(v->inactive_mask));
    if (v->unsupported_mask) parts.push_back("unsupported nvlinks=" + list(v->unsupported_mask));
    if (parts.empty()) return reason;
    std::string j;
    for (size_t i = 0; i < parts.size(); ++i) { if (i) j += "; "; j += parts[i]; }
    return reason + " (" + j + ")";
  };
  char buf[256];
  std::string o;
  switch (v->nvlink_reason) {
    case GPUD_NVLINK_NO_ISSUE: snprintf(buf, sizeof buf, "all %d GPU(s) were checked, no nvlink issue found", v->n_gpus); o = buf; break;
    case GPUD_NVLINK_P2P_FAILURE:
      snprintf(buf, sizeof buf, "no GPU pairs report NVLink P2P connectivity on %d-GPU NVLink-capable system", v->n_gpus); o = details(buf); break;
    case GPUD_NVLINK_NO_ACTIVE_LINKS:
      snprintf(buf, sizeof buf, "no GPUs report active nvlink links on %d-GPU NVLink-capable system", v->n_gpus); o = details(buf); break;
    case GPUD_NVLINK_THRESHOLD_SATISFIED:
      snprintf(buf, sizeof buf, "nvlink threshold satisfied: require >=%d GPUs with all links active; got %d", v->required, v->active); o = buf; break;
    case GPUD_NVLINK_THRESHOLD_VIOLATED:
      snprintf(buf, sizeof buf, "nvlink threshold violated: require >=%d GPUs with all links active; got %d", v->required, v->active); o = details(buf); break;
    case GPUD_NVLINK_NO_THRESHOLD: o = "nvlink threshold not set (skipped evaluation)"; break;
    case GPUD_NVLINK_NO_DATA: o = "no nvlink data (skipped evaluation)"; break;
    default: return -1;
  }
  if ((int)o.size() + 1 > cap) return -1;
  memcpy(out, o.c_str(), o.size() + 1);
  return (int32_t)o.size();
}

// setNVLinkSuggestedActions (nvlink/evaluate_threshold.go:37-52) + peerNVLinkStatusesSuggestReboot (component.go:398-415): does an
// unhealthy verdict come with RepairActionTypeRebootSystem?
extern "C" int32_t gpud_fabric_suggest_reboot(const gpud_fabric_verdict* v) {
  if (!v || v->nvlink_health != 2) return 0;
  const bool complete = v->p2p_expected_pairs != 0 && v->p2p_probed_pairs == v->p2p_expected_pairs;
  const uint32_t other = v->p2p_observed_status_mask & ~0x3eu;       // anything but the five "not supported" status codes
  return (v->inactive > 0 || (complete && v->p2p_ok_pairs == 0 && other != 0)) ? 1 : 0;
}

// nvml.Return.Error() is nvmlErrorString once go-nvml has loaded libnvidia-ml and a table of constant names before that; the text shows
// up in "status=..." issues.  NULL (the default) = the constant names.
static gpud_nvml_error_string_fn g_nvml_error_string = nullptr;
extern "C" void gpud_set_nvml_error_string(gpud_nvml_error_string_fn fn) { g_nvml_error_string = fn; }

// FabricState.GetIssues (pkg/nvidia/nvml/device/fabric_state.go:115-177) for one GPU's record: the sorted issue strings,
// joined with ", " (how fabric-manager/fabric_state.go:95-105 prints them); "" when healthy or when no fabric info was read.
extern "C" int32_t gpud_fabric_issues(const gpud_fabric_raw* g, char* out, int32_t cap) {
  if (!g || !out || cap <= 0) return -1;
  std::vector<std::string> is;
  if (g->fabric_valid) {
    if (g->fabric_state != 3) {
      static const char* kState[] = {"Not Supported", "Not Started", "In Progress", "Completed"};
      is.push_back(std::string("state=") + (g->fabric_state < 4 ? kState[g->fabric_state] : ("Unknown(" + std::to_string(g->fabric_state) + ")").c_str()));
    }
    if (g->fabric_status != 0 && g_nvml_error_string) {                       // FabricStatusToString = status.Error() (fabric_state.go:196-201)
      const char* txt = g_nvml_error_string(g->fabric_status);
      is.push_back(std::string("status=") + (txt ? txt : ""));
    } else if (g->fabric_status != 0) {
      const char* n = nullptr;
      switch (g->fabric_status) {            // go-nvml's built-in names, what Return.Error() gives while libnvidia-ml is not loaded (v0.13.0-1)
        case 1: n = "ERROR_UNINITIALIZED"; break; case 2: n = "ERROR_INVALID_ARGUMENT"; break; case 3: n = "ERROR_NOT_SUPPORTED"; break;
        case 4: n = "ERROR_NO_PERMISSION"; break; case 6: n = "ERROR_NOT_FOUND"; break; case 9: n = "ERROR_DRIVER_NOT_LOADED"; break;
        case 10: n = "ERROR_TIMEOUT"; break; case 15: n = "ERROR_GPU_IS_LOST"; break; case 999: n = "ERROR_UNKNOWN"; break;
      }
      is.push_back("status=" + (n ? std::string(n) : "ERROR_" + std::to_string(g->fabric_status)));
    }
    if (g->fabric_summary == 2) is.push_back("summary=Unhealthy");
    else if (g->fabric_summary == 3) is.push_back("summary=Limited Capacity");
    static const char* kMask[] = {"bandwidth degraded", "route recovery in progress", "route unhealthy", "access timeout recovery in progress"};
    for (int f = 0; f < 4; ++f)
      if (((g->fabric_health_mask >> (2 * f)) & 3u) == 1u) is.push_back(kMask[f]);
    std::sort(is.begin(), is.end());
  }
  std::string j;
  for (size_t i = 0; i < is.size(); ++i) { if (i) j += ", "; j += is[i]; }
  if ((int32_t)j.size() + 1 > cap) return -1;
  memcpy(out, j.c_str(), j.size() + 1);
  return (int32_t)j.size();
}

extern "C" int32_t gpud_ib_reason(const char* device, uint32_t port, int64_t t, int32_t flap, char* out, int32_t cap) {
  if (!device || !out || cap <= 0) return -1;
  // civil-from-days (proleptic Gregorian), UTC: time.Time.UTC().Format(time.RFC3339)
  int64_t days = t / 86400, rem = t % 86400;
  if (rem < 0) { rem += 86400; --days; }
  const int64_t z = days + 719468, era = (z >= 0 ? z : z - 146096) / 146097;
  const unsigned doe = (unsigned)(z - era * 146097), yoe = (doe - doe / 1460 + doe / 36524 - doe / 146096) / 365;
  const int64_t y = (int64_t)yoe + era * 400;
  const unsigned doy = doe - (365 * yoe + yoe / 4 - yoe / 100), mp = (5 * doy + 2) / 153, d = doy - (153 * mp + 2) / 5 + 1, m = mp < 10 ? mp + 3 : mp - 9;
  const int n = snprintf(out, (size_t)cap, "%s port %u down since %04lld-%02u-%02uT%02d:%02d:%02dZ%s", device, port, (long long)(y + (m <= 2)), m, d,
                         (int)(rem / 3600), (int)(rem % 3600 / 60), (int)(rem % 60), flap ? " (and flapped back to active)" : "");
  return n < cap ? n : -1;
}

// xidErrorEventDetail JSON (xid/health_state.go:284-315), field order and omitempty as encoding/json emits them;
// time is RFC3339 UTC seconds like metav1.Time.
extern "C" int32_t gpud_hit_detail_json(const gpud_xid_hit* h, int64_t unix_seconds, char* out, int32_t cap) {
  if (!h || !out || cap < 2) return GPUD_E_INVALID;
  if (h->kind == GPUD_KIND_SXID) {   // sxid events persist only the decimal code (sxid/component.go, health_state.go:113-142)
    const int n = snprintf(out, (size_t)cap, "%d", h->code);
    return n < cap ? GPUD_OK : GPUD_E_CAPACITY;
  }
  std::string o = "{\"time\":";
  if (unix_seconds == 0) {
    o += "null";
  } else {
    time_t t = (time_t)unix_seconds;
    struct tm tmv;
    gmtime_r(&t, &tmv);
    char tb[40];
    strftime(tb, sizeof tb, "\"%Y-%m-%dT%H:%M:%SZ\"", &tmv);
    o += tb;
  }
  o += ",\"data_source\":\"kmsg\",\"device_uuid\":\"";
  json_escape(o, h->device);
  o += "\",\"xid\":" + std::to_string(h->code);
  const bool ext = (h->flags & GPUD_HIT_EXTENDED) != 0;
  if (ext && h->sub_code != 0) o += ",\"sub_code\":" + std::to_string(h->sub_code);
  if (ext && h->unit_name[0]) { o += ",\"sub_code_description\":\""; json_escape(o, h->unit_name); o += "\""; }
  if (ext && h->error_status != 0) o += ",\"error_status\":" + std::to_string(h->error_status);
  const char* hint = (h->flags & GPUD_HIT_HAS_RULE) ? gpud_nvlink_rule_hint(h->rule_index) : "";
  if (hint[0]) { o += ",\"investigatory_hint\":\""; json_escape(o, hint); o += "\""; }
  const char* desc = gpud_xid_description(h->code, h->detail_variant);
  if (desc[0]) { o += ",\"description\":\""; json_escape(o, desc); o += "\""; }
  if (h->n_actions >= 0) {
    o += ",\"suggested_actions_by_gpud\":{\"description\":\"\",\"repair_actions\":[";   // apiv1.SuggestedActions: description is not omitempty (api/v1/types.go:206-212)
    for (int i = 0; i < h->n_actions; ++i) {
      if (i) o += ",";
      o += "\"";
      o += action_wire(h->actions[i]);
      o += "\"";
    }
    o += "]}";
  }
  o += "}";
  if ((int)o.size() + 1 > cap) return GPUD_E_CAPACITY;
  memcpy(out, o.c_str(), o.size() + 1);
  return GPUD_OK;
}
