// host_component.cpp — see host_component.h.  Every function cites the reference code it mirrors.
#include "host_component.h"
#include "json_min.h"

#include <string.h>
#include <time.h>

#include <algorithm>

namespace gpud {

const char* health_string(Health h) {            // api/v1/types.go:20-25
  switch (h) {
    case Health::Healthy: return "Healthy";
    case Health::Degraded: return "Degraded";
    case Health::Unhealthy: return "Unhealthy";
    case Health::Initializing: return "Initializing";
  }
  return "Healthy";
}
const char* event_type_string(int32_t t) {       // api/v1/types.go:222-244
  static const char* n[] = {"Unknown", "Info", "Warning", "Critical", "Fatal"};
  return (t >= 0 && t <= 4) ? n[t] : "Unknown";
}
const char* repair_action_string(int32_t a) {    // api/v1/types.go:183-203
  switch (a) {
    case GPUD_ACT_IGNORE_NO_ACTION_REQUIRED: return "IGNORE_NO_ACTION_REQUIRED";
    case GPUD_ACT_REBOOT_SYSTEM: return "REBOOT_SYSTEM";
    case GPUD_ACT_HARDWARE_INSPECTION: return "HARDWARE_INSPECTION";
    case GPUD_ACT_CHECK_USER_APP_AND_GPU: return "CHECK_USER_APP_AND_GPU";
  }
  return "";
}

void jstr(std::string& o, const std::string& s) {
  o.push_back('"');
  for (unsigned char c : s) {
    if (c == '"') o += "\\\"";
    else if (c == '\\') o += "\\\\";
    else if (c == '\n') o += "\\n";
    else if (c < 0x20 || c == '<' || c == '>' || c == '&') { char b[8]; snprintf(b, sizeof b, "\\u%04x", c); o += b; }
    else o.push_back((char)c);
  }
  o.push_back('"');
}

std::string HealthState::to_json() const {       // api/v1/types.go:50-94 (omitempty fields)
  std::string o = "{";
  bool first = true;
  auto key = [&](const char* k) { if (!first) o += ","; first = false; o += "\""; o += k; o += "\":"; };
  key("time");                                   // metav1.Time: RFC3339 seconds in UTC, null when zero
  if (time_unix == 0) o += "null";
  else {
    time_t t = (time_t)time_unix;
    struct tm tmv;
    gmtime_r(&t, &tmv);
    char tb[40];
    strftime(tb, sizeof tb, "\"%Y-%m-%dT%H:%M:%SZ\"", &tmv);
    o += tb;
  }
  if (!component.empty()) { key("component"); jstr(o, component); }
  if (!name.empty()) { key("name"); jstr(o, name); }
  key("health"); jstr(o, health_string(health));
  if (!reason.empty()) { key("reason"); jstr(o, reason); }
  if (!error.empty()) { key("error"); jstr(o, error); }
  if (has_actions) {
    key("suggested_actions");
    o += "{\"description\":\"\",\"repair_actions\":[";    // neither field of apiv1.SuggestedActions is omitempty (types.go:206-212)
    for (size_t i = 0; i < actions.repair_actions.size(); ++i) { if (i) o += ","; jstr(o, repair_action_string(actions.repair_actions[i])); }
    o += "]}";
  }
  o += "}";
  return o;
}

// ---- pkg/kmsg ----
static bool go_atoi64(const std::string& s, int64_t* out) {   // strconv.Atoi / ParseInt(s, 10, 64)
  size_t i = 0;
  bool neg = false;
  if (!s.empty() && (s[0] == '+' || s[0] == '-')) { neg = s[0] == '-'; i = 1; }
  if (i >= s.size()) return false;
  unsigned long long v = 0;
  const unsigned long long lim = neg ? 0x8000000000000000ull : 0x7fffffffffffffffull;
  for (; i < s.size(); ++i) {
    if (s[i] < '0' || s[i] > '9') return false;
    const unsigned d = (unsigned)(s[i] - '0');
    if (v > (lim - d) / 10ull) return false;
    v = v * 10ull + d;
  }
  *out = neg ? (int64_t)(0ull - v) : (int64_t)v;
  return true;
}

bool parse_kmsg_line(const std::string& line, KmsgMessage* out, std::string* err) {   // watcher.go:292-332
  const size_t semi = line.find(';');
  if (semi == std::string::npos) { if (err) *err = "invalid kmsg; must contain a ';'"; return false; }
  const std::string meta = line.substr(0, semi);
  std::vector<std::string> parts;
  size_t b = 0;
  for (;;) {
    const size_t c = meta.find(',', b);
    parts.push_back(meta.substr(b, c == std::string::npos ? std::string::npos : c - b));
    if (c == std::string::npos) break;
    b = c + 1;
  }
  if (parts.size() < 3) { if (err) *err = "invalid kmsg: must contain at least 3 ',' separated pieces at the start"; return false; }
  int64_t prio, seq, usec;
  if (!go_atoi64(parts[0], &prio)) { if (err) *err = "could not parse priority"; return false; }
  if (!go_atoi64(parts[1], &seq)) { if (err) *err = "could not parse sequence number"; return false; }
  if (!go_atoi64(parts[2], &usec)) { if (err) *err = "could not parse timestamp"; return false; }
  out->priority = (int)prio; out->sequence = seq; out->usec_since_boot = usec;
  out->message = line.substr(semi + 1);            // unmodified: continuation lines and extra ';' stay (watcher_test.go:113-133)
  return true;
}

std::string dedup_key(int64_t unix_seconds, const std::string& message, int truncate_seconds) {   // deduper.go:63-74
  if (truncate_seconds <= 0) truncate_seconds = 60;
  const int64_t t = unix_seconds - (unix_seconds % truncate_seconds);
  return std::to_string(t) + "-" + message;
}

int Deduper::add(int64_t now_unix, int64_t msg_unix, const std::string& message) {      // deduper.go:111-125
  const std::string k = dedup_key(msg_unix, message, trunc_);
  auto it = cache_.find(k);
  int freq = 1;
  if (it != cache_.end() && it->second.expires >= now_unix) freq = it->second.count + 1;      // go-cache: an item is gone once now > its expiration
  cache_[k] = Entry{freq, now_unix + ttl_};
  return freq;
}

// ---- eventstore ----
static bool same_event(const Event& a, const Event& b) {       // findEvent's WHERE + compareEvent (database.go:278-324, 459-469)
  if (a.time_unix != b.time_unix || a.name != b.name || a.type != b.type) return false;
  if (!b.message.empty() && a.message != b.message) return false;
  return a.extra_info == b.extra_info;
}
bool EventBucket::insert(const Event& ev) { events_.push_back(ev); return true; }
const Event* EventBucket::find(const Event& ev) const {
  for (const Event& e : events_) if (same_event(e, ev)) return &e;
  return nullptr;
}
std::vector<Event> EventBucket::get(int64_t since_unix) const {
  std::vector<Event> out;
  for (const Event& e : events_) if (e.time_unix > since_unix) out.push_back(e);   // getEvents: WHERE timestamp > ? (pkg/eventstore/database.go:330)
  std::stable_sort(out.begin(), out.end(), [](const Event& a, const Event& b) { return a.time_unix > b.time_unix; });
  return out;
}
int EventBucket::purge(int64_t before_unix) {
  const size_t n0 = events_.size();
  events_.erase(std::remove_if(events_.begin(), events_.end(), [&](const Event& e) { return e.time_unix < before_unix; }), events_.end());
  return (int)(n0 - events_.size());
}

// ---- health evolution ----
EvolveResult evolve_healthy_state(const std::vector<XidEventView>& ev, int reboot_threshold, const char* event_name) {   // xid/health_state.go:57-128 ; sxid/health_state.go:38-111 (name "error_sxid", threshold 2)
  EvolveResult r;
  int last_health = 0;
  std::map<uint64_t, int> reboot_map;
  for (auto it = ev.rbegin(); it != ev.rend(); ++it) {          // oldest -> newest
    if (it->name == event_name) {
      int cur = 0;
      if (it->type == "Critical") cur = 1;
      else if (it->type == "Fatal") cur = 2;
      if (cur < last_health) continue;
      last_health = cur;
      r.has_xid = true;
      r.xid = it->xid;
      r.last_index = (int)(ev.rend() - it) - 1;
      if (it->has_actions && !it->actions.empty()) {
        std::vector<int32_t> acts = it->actions;
        if (acts[0] == GPUD_ACT_REBOOT_SYSTEM) {
          auto f = reboot_map.find(it->xid);
          if (f == reboot_map.end()) reboot_map[it->xid] = 0;
          else if (f->second >= reboot_threshold) acts[0] = GPUD_ACT_HARDWARE_INSPECTION;
        }
        acts.resize(1);
        r.has_actions = true;
        r.actions = acts;
      }
    } else if (it->name == "reboot") {
      if (r.has_actions && !r.actions.empty() && (r.actions[0] == GPUD_ACT_REBOOT_SYSTEM || r.actions[0] == GPUD_ACT_CHECK_USER_APP_AND_GPU)) {
        last_health = 0;
        r.has_actions = false;
        r.actions.clear();
        r.has_xid = false;
        r.xid = 0;
        r.last_index = -1;
      }
      for (auto& kv : reboot_map) kv.second += 1;
    }
  }
  r.health = last_health == 2 ? Health::Unhealthy : (last_health == 1 ? Health::Degraded : Health::Healthy);
  return r;
}

// ---- stored xid events: payload, resolve, message ----
static bool parse_actions_object(const char*& p, XidPayload* out) {      // {"repair_actions":[...]} | null
  jsonmin::ws(p);
  if (!strncmp(p, "null", 4)) { p += 4; return true; }
  if (*p != '{') return false;
  ++p;
  out->has_actions = true;
  jsonmin::ws(p);
  if (*p == '}') { ++p; return true; }
  for (;;) {
    std::string k;
    jsonmin::ws(p);
    if (!jsonmin::string(p, &k)) return false;
    jsonmin::ws(p);
    if (*p++ != ':') return false;
    jsonmin::ws(p);
    if (k == "repair_actions" && *p == '[') {
      ++p;
      jsonmin::ws(p);
      if (*p == ']') ++p;
      else for (;;) {
        std::string a;
        jsonmin::ws(p);
        if (!jsonmin::string(p, &a)) return false;
        int id = 0;                                                    // an unknown action string stays in the list as itself in Go; 0 here
        for (int i = 1; i <= 4; ++i) if (a == repair_action_string(i)) id = i;
        out->actions.push_back(id);
        jsonmin::ws(p);
        if (*p == ',') { ++p; continue; }
        if (*p == ']') { ++p; break; }
        return false;
      }
    } else if (!jsonmin::skip(p)) return false;
    jsonmin::ws(p);
    if (*p == ',') { ++p; continue; }
    if (*p == '}') { ++p; return true; }
    return false;
  }
}

static bool parse_uint(const char*& p, uint64_t* v) {
  if (*p < '0' || *p > '9') return false;
  uint64_t x = 0;
  while (*p >= '0' && *p <= '9') {
    if (x > (UINT64_MAX - (uint64_t)(*p - '0')) / 10) return false;
    x = x * 10 + (uint64_t)(*p++ - '0');
  }
  if (*p == '.' || *p == 'e' || *p == 'E') return false;               // encoding/json refuses a non-integer for an integer field
  *v = x;
  return true;
}

static bool parse_xid_payload(const std::string& raw, XidPayload* out) {   // json.Unmarshal([]byte(rawData), &xidErr) == nil
  *out = XidPayload();
  const char* p = raw.c_str();
  jsonmin::ws(p);
  if (*p != '{') return false;
  ++p;
  jsonmin::ws(p);
  if (*p == '}') { ++p; jsonmin::ws(p); return *p == 0; }
  for (;;) {
    std::string k;
    jsonmin::ws(p);
    if (!jsonmin::string(p, &k)) return false;
    jsonmin::ws(p);
    if (*p++ != ':') return false;
    jsonmin::ws(p);
    uint64_t u = 0;
    if (!strncmp(p, "null", 4) && k != "suggested_actions_by_gpud") p += 4;                 // null leaves the zero value
    else if (k == "xid") { if (!parse_uint(p, &u)) return false; out->xid = u; }
    else if (k == "sub_code") {
      const bool neg = *p == '-';
      if (neg) ++p;
      if (!parse_uint(p, &u) || u > 0x7fffffffu) return false;
      out->sub_code = neg ? -(int32_t)u : (int32_t)u;
    }
    else if (k == "error_status") { if (!parse_uint(p, &u) || u > 0xffffffffull) return false; out->error_status = (uint32_t)u; }
    else if (k == "device_uuid") { if (!jsonmin::string(p, &out->device_uuid)) return false; }
    else if (k == "sub_code_description") { if (!jsonmin::string(p, &out->sub_code_description)) return false; }
    else if (k == "description") { if (!jsonmin::string(p, &out->description)) return false; }
    else if (k == "suggested_actions_by_gpud") { if (!parse_actions_object(p, out)) return false; }
    else if (!jsonmin::skip(p)) return false;                                              // time, data_source, investigatory_hint, unknown keys
    jsonmin::ws(p);
    if (*p == ',') { ++p; continue; }
    if (*p == '}') { ++p; break; }
    return false;
  }
  jsonmin::ws(p);
  return *p == 0;
}

std::string convert_bus_id_to_uuid(const std::string& bus_id, const DeviceMap& devices) {
  for (const auto& kv : devices)
    if (gpud_xid_device_matches_bus_id(bus_id.c_str(), kv.second.c_str())) return kv.first;
  return "";
}

std::string xid_payload_message(const XidPayload& p, const DeviceMap& devices) {
  char buf[1536];
  const std::string uuid = convert_bus_id_to_uuid(p.device_uuid, devices);
  const int n = gpud_xid_build_message(p.xid, p.sub_code, p.error_status, p.description.c_str(), p.device_uuid.c_str(), uuid.c_str(), buf, sizeof buf);
  return n >= 0 ? std::string(buf, (size_t)n) : std::string("unknown");
}

bool resolve_xid_event(std::string* type, const std::string& raw, const std::string& event_device_uuid, const DeviceMap& devices, XidPayload* out,
                       std::string* message) {
  XidPayload p;
  if (!parse_xid_payload(raw, &p) || p.xid == 0) {
    // legacy rows: only the decimal code (strconv.Atoi), the device in ExtraInfo["device_uuid"]
    const char* q = raw.c_str();
    bool neg = false;
    if (*q == '+' || *q == '-') { neg = *q == '-'; ++q; }
    uint64_t u = 0;
    const char* q0 = q;
    while (*q >= '0' && *q <= '9' && u < (1ull << 40)) u = u * 10 + (uint64_t)(*q++ - '0');
    if (q == q0 || *q != 0 || neg || u > 0x7fffffffu) return false;
    int32_t n_act = -1, acts[4] = {0, 0, 0, 0};
    if (!gpud_xid_get_detail((int32_t)u, nullptr, &n_act, acts)) return false;          // GetDetail(currXid), the base entry (:203)
    p = XidPayload();
    p.xid = u;
    p.device_uuid = event_device_uuid;
    p.has_actions = n_act >= 0;
    for (int i = 0; i < n_act; ++i) p.actions.push_back(acts[i]);
  }
  // addEventDetails (:222-281)
  if (p.xid <= (uint64_t)INT64_MAX) {
    int32_t ev = 0, n_act = -1, acts[4] = {0, 0, 0, 0}, variant = 0, sc = 0;
    const bool found = p.xid <= 0x7fffffffu && p.sub_code >= 0 && gpud_xid_detail((int32_t)p.xid, p.sub_code, p.error_status, &ev, &n_act, acts, &variant, &sc);
    if (found) {
      if (type->empty() && ev != GPUD_EVENT_UNKNOWN) *type = event_type_string(ev);
      if (p.description.empty()) p.description = gpud_xid_description((int32_t)p.xid, variant);
      if (p.sub_code == 0) p.sub_code = sc;
      if (!p.has_actions && n_act >= 0) { p.has_actions = true; p.actions.assign(acts, acts + n_act); }
      // SubCodeDescription is not part of any message or decision; the merged tables keep the unit only in the reference's Detail
    } else if (type->empty()) {
      *type = "Unknown";
    }
  } else {
    return false;                                                      // intFromUint64 fails: the event is returned unchanged
  }
  if (message) *message = xid_payload_message(p, devices);
  *out = p;
  return true;
}

StoredEvolveResult evolve_stored_events(const std::vector<Event>& events, const DeviceMap& devices, int reboot_threshold) {
  std::vector<XidEventView> views;
  std::vector<XidPayload> payloads;
  for (const Event& e : events) {
    XidEventView v;
    v.name = e.name;
    XidPayload p;
    if (e.name == "error_xid") {
      std::string type = e.type;
      auto it = e.extra_info.find("data");
      auto dv = e.extra_info.find("device_uuid");
      if (!resolve_xid_event(&type, it == e.extra_info.end() ? std::string() : it->second, dv == e.extra_info.end() ? std::string() : dv->second, devices, &p,
                             nullptr))
        continue;                                                      // unmarshal of the unresolved payload fails -> skipped (:71-74)
      v.type = type;
      v.xid = p.xid;
      v.has_actions = p.has_actions;
      v.actions = p.actions;
    }
    views.push_back(v);
    payloads.push_back(p);
  }
  const EvolveResult r = evolve_healthy_state(views, reboot_threshold, "error_xid");
  StoredEvolveResult out;
  out.health = r.health;
  out.has_actions = r.has_actions;
  out.actions = r.actions;
  out.reason = (r.has_xid && r.last_index >= 0) ? xid_payload_message(payloads[(size_t)r.last_index], devices) : "XIDComponent is healthy";
  return out;
}

// ---- stored sxid events ----
static bool parse_sxid_payload(const std::string& raw, SXidPayload* out) {     // json.Unmarshal into sxidErrorEventDetail (health_state.go:144-162)
  *out = SXidPayload();
  const char* p = raw.c_str();
  jsonmin::ws(p);
  if (*p != '{') return false;
  ++p;
  jsonmin::ws(p);
  if (*p == '}') { ++p; jsonmin::ws(p); return *p == 0; }
  for (;;) {
    std::string k;
    jsonmin::ws(p);
    if (!jsonmin::string(p, &k)) return false;
    jsonmin::ws(p);
    if (*p++ != ':') return false;
    jsonmin::ws(p);
    uint64_t u = 0;
    XidPayload tmp;
    if (!strncmp(p, "null", 4) && k != "suggested_actions_by_gpud") p += 4;
    else if (k == "sxid") { if (!parse_uint(p, &u)) return false; out->sxid = u; }
    else if (k == "device_uuid") { if (!jsonmin::string(p, &out->device_uuid)) return false; }
    else if (k == "suggested_actions_by_gpud") {
      if (!parse_actions_object(p, &tmp)) return false;
      out->has_actions = tmp.has_actions;
      out->actions = tmp.actions;
    }
    else if (!jsonmin::skip(p)) return false;
    jsonmin::ws(p);
    if (*p == ',') { ++p; continue; }
    if (*p == '}') { ++p; break; }
    return false;
  }
  jsonmin::ws(p);
  return *p == 0;
}

// resolveSXIDEvent (sxid/health_state.go:113-142) followed by the Unmarshal of evolveHealthyState (:50-54): false = skipped
bool resolve_sxid_event(std::string* type, const std::string& raw, const std::string& event_device_uuid, SXidPayload* out, std::string* message) {
  const char* q = raw.c_str();
  bool neg = false;
  if (*q == '+' || *q == '-') { neg = *q == '-'; ++q; }
  const char* q0 = q;
  uint64_t u = 0;
  while (*q >= '0' && *q <= '9' && u < (1ull << 40)) u = u * 10 + (uint64_t)(*q++ - '0');
  if (q != q0 && *q == 0) {                                            // strconv.Atoi succeeded
    if (u > 0x7fffffffu) return false;                                 // no such catalog entry: event unchanged, its decimal payload does not unmarshal
    int32_t ev = 0, n_act = -1, acts[4] = {0, 0, 0, 0};
    if (!gpud_sxid_get_detail(neg ? -(int32_t)u : (int32_t)u, &ev, &n_act, acts)) return false;
    *type = event_type_string(ev);
    if (neg) return false;                                             // uint64FromInt fails: type / message set, payload still the decimal string
    out->sxid = u;
    out->device_uuid = event_device_uuid;
    out->has_actions = n_act >= 0;
    out->actions.assign(acts, acts + (n_act > 0 ? n_act : 0));
    if (message) {
      char buf[512];
      const int n = gpud_sxid_reason((int64_t)u, event_device_uuid.c_str(), buf, sizeof buf);
      *message = n >= 0 ? std::string(buf, (size_t)n) : std::string();
    }
    return true;
  }
  return parse_sxid_payload(raw, out);                                 // already the JSON form: type stays the event's
}

StoredEvolveResult evolve_stored_sxid_events(const std::vector<Event>& events) {
  std::vector<XidEventView> views;
  std::vector<SXidPayload> payloads;
  for (const Event& e : events) {
    XidEventView v;
    v.name = e.name;
    SXidPayload p;
    if (e.name == "error_sxid") {
      std::string type = e.type;
      auto it = e.extra_info.find("data");
      auto dv = e.extra_info.find("device_uuid");
      if (!resolve_sxid_event(&type, it == e.extra_info.end() ? std::string() : it->second, dv == e.extra_info.end() ? std::string() : dv->second, &p, nullptr))
        continue;
      v.type = type;
      v.xid = p.sxid;
      v.has_actions = p.has_actions;
      v.actions = p.actions;
    }
    views.push_back(v);
    payloads.push_back(p);
  }
  const EvolveResult r = evolve_healthy_state(views, 2 /* rebootThreshold, sxid/health_state.go:36 */, "error_sxid");
  StoredEvolveResult out;
  out.health = r.health;
  out.has_actions = r.has_actions;
  out.actions = r.actions;
  char buf[512];
  int n;
  if (r.has_xid && r.last_index >= 0) {
    const SXidPayload& p = payloads[(size_t)r.last_index];
    n = p.sxid <= (uint64_t)INT64_MAX ? gpud_sxid_reason((int64_t)p.sxid, p.device_uuid.c_str(), buf, sizeof buf)
                                      : snprintf(buf, sizeof buf, "SXID %llu detected on %s", (unsigned long long)p.sxid, p.device_uuid.c_str());
  } else {
    n = gpud_sxid_reason(-1, "", buf, sizeof buf);
  }
  out.reason = n >= 0 ? std::string(buf, (size_t)n) : std::string();
  return out;
}

std::vector<Event> trim_events_after_set_healthy(const std::vector<Event>& e) {   // component.go:630-642
  for (size_t i = 0; i < e.size(); ++i)
    if (e[i].name == "SetHealthy") return std::vector<Event>(e.begin(), e.begin() + i);
  return e;
}
std::vector<Event> merge_events(const std::vector<Event>& a, const std::vector<Event>& b) {   // component.go:614-628
  std::vector<Event> r(a);
  r.insert(r.end(), b.begin(), b.end());
  std::stable_sort(r.begin(), r.end(), [](const Event& x, const Event& y) { return x.time_unix > y.time_unix; });
  return r;
}

// ---- threshold rules ----
// time.Duration.String() for a whole number of seconds ("10m0s", "1h0m0s", "45s", "0s", "-10m0s"): what %s prints for the window.
std::string go_duration_seconds(int64_t sec) {
  if (sec == 0) return "0s";
  std::string out = sec < 0 ? "-" : "";
  uint64_t u = sec < 0 ? (uint64_t)(-(sec + 1)) + 1 : (uint64_t)sec;
  const uint64_t h = u / 3600, m = (u / 60) % 60, s = u % 60;
  if (h) out += std::to_string(h) + "h";
  if (h || m) out += std::to_string(m) + "m";
  return out + std::to_string(s) + "s";
}

SlowdownVerdict evaluate_hw_slowdown(const std::vector<int64_t>& ev, int64_t now, int64_t window_s, double thr) {   // hw-slowdown/component.go:352-407
  SlowdownVerdict v;
  if (window_s == 0) { v.reason = "no time window to evaluate states"; return v; }
  const int64_t since = now - window_s;
  std::vector<int64_t> minutes;
  for (int64_t t : ev) if (t > since) minutes.push_back(t / 60);         // eventBucket.Get(since) is "timestamp > since" (eventstore/database.go:330), then Unix()/60
  if (minutes.empty()) { v.reason = "no clock events found"; return v; }
  std::sort(minutes.begin(), minutes.end());
  minutes.erase(std::unique(minutes.begin(), minutes.end()), minutes.end());
  v.distinct_minutes = (int)minutes.size();
  v.freq_per_min = (double)v.distinct_minutes / ((double)window_s / 60.0);
  char buf[320];
  const std::string win = go_duration_seconds(window_s);
  if (v.freq_per_min < thr) {
    snprintf(buf, sizeof buf, "hw slowdown events frequency per minute %.2f (total events per minute count %d) is less than threshold %.2f for the last %s",
             v.freq_per_min, v.distinct_minutes, thr, win.c_str());
    v.reason = buf;
    return v;
  }
  v.health = Health::Unhealthy;
  v.inspect = true;
  snprintf(buf, sizeof buf, "hw slowdown events frequency per minute %.2f (total events per minute count %d) exceeded threshold %.2f for the last %s",
           v.freq_per_min, v.distinct_minutes, thr, win.c_str());
  v.reason = buf;
  return v;
}

// The per-GPU rules of temperature Check (temperature/component.go:206-248).  Returns a bit set: 1 GPU core above its max-operating
// threshold, 2 HBM above the memory max, 4 thermal margin at or below the configured margin threshold.  The reason the component
// reports is the first of margin > gpu > hbm (:273-287).
int evaluate_temperature(const TemperatureReading& t, int32_t margin_threshold) {
  int m = 0;
  if (t.threshold_gpu_max > 0 && t.current_gpu_core > t.threshold_gpu_max) m |= 1;          // strict '>' (the compare n_over uses too)
  if (t.threshold_mem_max > 0 && t.hbm_supported && t.current_hbm > t.threshold_mem_max) m |= 2;
  if (t.threshold_slowdown > 0 && t.margin_supported && margin_threshold > 0 && t.slowdown_margin > 0 && t.slowdown_margin <= margin_threshold) m |= 4;
  return m;
}

// ---- the component ----
XidComponent::XidComponent(gpud_ctx* ctx, int32_t dev, bool row_remap, int reboot_threshold)
    : ctx_(ctx), dev_(dev), row_remap_(row_remap), reboot_threshold_(reboot_threshold) {
  cur_.name = "error_xid";                         // StateNameErrorXid, component.go:39
  cur_.component = kName;
  cur_.health = Health::Healthy;
  cur_.reason = "XIDComponent is healthy";
}
void XidComponent::AddRebootEvent(int64_t t) {
  Event e;
  e.time_unix = t;
  e.name = "reboot";
  std::lock_guard<std::mutex> g(mu_);
  reboots_.insert(e);
}
int32_t XidComponent::Start() {
  std::lock_guard<std::mutex> g(mu_);
  update_state();
  return 0;
}

static bool keep_hit(const gpud_xid_hit& h, bool row_remap) {
  // Xid 63/64 are left to the remapped-rows component when the product supports row remapping (component.go:290, 484)
  return !(h.kind == GPUD_KIND_XID && row_remap && (h.code == 63 || h.code == 64));
}

CheckResult XidComponent::Check() {               // component.go:255-311
  CheckResult cr;
  cr.component = kName;
  std::vector<gpud_xid_hit> hits(4096);
  int64_t n_hits = 0, n_units = 0;
  for (;;) {
    const int32_t rc = gpud_kmsg_scan(ctx_, dev_, reinterpret_cast<const uint8_t*>(buf_.data()), (int64_t)buf_.size(),
                                      raw_ ? GPUD_SCAN_RAW_KMSG : GPUD_SCAN_LINES, hits.data(), (int64_t)hits.size(), &n_hits, &n_units);
    if (rc == GPUD_E_CAPACITY && n_hits > (int64_t)hits.size()) { hits.resize((size_t)n_hits); continue; }
    if (rc != GPUD_OK) {
      char msg[512] = {0};
      gpud_last_error(ctx_, msg, sizeof msg);
      cr.health = Health::Unhealthy;              // errors are embedded in the result, Check never throws (types.go:48-53)
      cr.summary = "failed to read kmsg";
      HealthState s = cur_;
      s.health = Health::Unhealthy; s.reason = cr.summary; s.error = msg;
      cr.states.push_back(s);
      return cr;
    }
    break;
  }
  for (int64_t i = 0; i < n_hits; ++i)
    if (hits[i].kind == GPUD_KIND_XID && keep_hit(hits[i], row_remap_)) cr.found.push_back(hits[i]);
  cr.summary = "matched " + std::to_string(cr.found.size()) + " xid errors from " + std::to_string(n_units) + " kmsg(s)";
  cr.health = Health::Healthy;
  for (const auto& h : cr.found)
    if (h.event_type == GPUD_EVENT_CRITICAL || h.event_type == GPUD_EVENT_FATAL) { cr.health = Health::Unhealthy; break; }
  HealthState s;
  s.name = "error_xid"; s.component = kName; s.health = cr.health; s.reason = cr.summary;
  cr.states.push_back(s);
  cr.text = cr.found.empty() ? "no xid error found" : cr.summary;
  std::lock_guard<std::mutex> g(mu_);
  checked_ = true;
  return cr;
}

int32_t XidComponent::IngestHits(const std::vector<gpud_xid_hit>& hits, int64_t fallback_unix) {   // component.go:468-577
  std::lock_guard<std::mutex> g(mu_);
  int inserted = 0;
  for (const auto& h : hits) {
    if (h.kind != GPUD_KIND_XID || !keep_hit(h, row_remap_)) continue;
    const int64_t t = raw_ ? boot_unix_ + h.kmsg_usec / 1000000 : fallback_unix;
    char payload[4096];
    if (gpud_hit_detail_json(&h, t, payload, sizeof payload) != GPUD_OK) continue;
    Event ev;
    ev.time_unix = t;
    ev.name = "error_xid";
    ev.type = event_type_string(h.event_type);
    ev.extra_info["device_uuid"] = h.device;
    ev.extra_info["data"] = payload;
    if (bucket_.find(ev)) continue;                // "find the same event, skip inserting it"
    bucket_.insert(ev);
    ++inserted;
  }
  if (inserted) update_state();
  return inserted;
}

void XidComponent::update_state() {                // component.go:581-611
  std::vector<Event> local = trim_events_after_set_healthy(bucket_.get(INT64_MIN));
  std::vector<Event> all = merge_events(reboots_.get(INT64_MIN), local);
  const StoredEvolveResult r = evolve_stored_events(all, devices_, reboot_threshold_);
  cur_.health = r.health;
  cur_.has_actions = r.has_actions;
  cur_.actions.repair_actions = r.actions;
  cur_.reason = r.reason;
}

std::vector<HealthState> XidComponent::LastHealthStates() {     // types.go:55-58
  std::lock_guard<std::mutex> g(mu_);
  return {cur_};
}
std::vector<Event> XidComponent::Events(int64_t since) {
  std::lock_guard<std::mutex> g(mu_);
  return bucket_.get(since);
}
int32_t XidComponent::SetHealthy(int64_t now_unix) {            // set_healthy.go:14-35
  std::lock_guard<std::mutex> g(mu_);
  Event e;
  e.time_unix = now_unix;
  e.name = "SetHealthy";
  bucket_.insert(e);
  update_state();
  return 0;
}

}  // namespace gpud

// getClockEventReasons (hw-slowdown/clock_events.go:168-190) over clockEventReasonsToInclude (:192-264): the descriptions of the
// set bits of an nvmlDeviceGetCurrentClocksEventReasons bitmask, HW-slowdown ones and the others, each sorted; one per line.
namespace {
struct ClockReason { unsigned long long flag; int hw; const char* text; };
const ClockReason kClockReasons[] = {
    {0x1ull, 0, "GPU is idle and clocks are dropping to Idle state"},
    {0x2ull, 0, "GPU clocks are limited by current setting of applications clocks"},
    {0x4ull, 0, "Clocks have been optimized to not exceed currently set power limits ('SW Power Cap: Active' in nvidia-smi --query)"},
    {0x8ull, 1, "HW Slowdown is engaged due to high temperature, power brake assertion, or high power draw ('HW Slowdown: Active' in nvidia-smi --query)"},
    {0x10ull, 0, "GPU is part of a Sync boost group to maximize performance per watt"},
    {0x20ull, 0, "SW Thermal Slowdown is active to keep GPU and memory temperatures within operating limits"},
    {0x40ull, 1, "HW Thermal Slowdown (reducing the core clocks by a factor of 2 or more) is engaged (temperature being too high) ('HW Thermal Slowdown' in nvidia-smi --query)"},
    {0x80ull, 1, "HW Power Brake Slowdown (reducing the core clocks by a factor of 2 or more) is engaged (External Power Brake Assertion being triggered) ('HW Power Brake Slowdown' in nvidia-smi --query)"},
    {0x100ull, 0, "GPU clocks are limited by current setting of Display clocks"},
};
}  // namespace
extern "C" int32_t gpud_clock_event_reasons(uint64_t bitmask, char* hw_out, int32_t hw_cap, char* other_out, int32_t other_cap, int32_t* flags3) {
  std::vector<std::string> hw, other;
  for (const ClockReason& r : kClockReasons)
    if (bitmask & r.flag) (r.hw ? hw : other).push_back(r.text);
  std::sort(hw.begin(), hw.end());
  std::sort(other.begin(), other.end());
  auto join = [](const std::vector<std::string>& v, char* out, int32_t cap) -> bool {
    std::string j;
    for (size_t i = 0; i < v.size(); ++i) { if (i) j += "\n"; j += v[i]; }
    if (!out || (int32_t)j.size() + 1 > cap) return false;
    memcpy(out, j.c_str(), j.size() + 1);
    return true;
  };
  if (flags3) {      // ClockEvents.HWSlowdown / HWSlowdownThermal / HWSlowdownPowerBrake (clock_events.go:151-153)
    flags3[0] = (bitmask & 0x8ull) != 0; flags3[1] = (bitmask & 0x40ull) != 0; flags3[2] = (bitmask & 0x80ull) != 0;
  }
  return (join(hw, hw_out, hw_cap) && join(other, other_out, other_cap)) ? (int32_t)(hw.size() * 100 + other.size()) : -1;
}

// The temperature component's check result over the box's readings (temperature/component.go:190-287): health 0 Healthy / 1 Degraded
// and the reason -- the per-GPU findings of the first non-empty class (margin, GPU, HBM) joined by ", ", in the order given (the
// reference ranges over a Go map, so its order is unspecified).
extern "C" int32_t gpud_temperature_reason(const gpud_temperature* ts, const char* const* gpu_uuids, int32_t n, int32_t margin_threshold_c, int32_t* health,
                                           char* out, int32_t cap) {
  if (n < 0 || (n && (!ts || !gpu_uuids)) || !out || cap <= 0) return -1;
  std::vector<std::string> margin, gpu, hbm;
  char buf[384];
  for (int32_t i = 0; i < n; ++i) {
    int32_t bits = 0;
    if (gpud_temperature_check(&ts[i], margin_threshold_c, &bits) != GPUD_OK) return -1;
    const char* uuid = gpu_uuids[i] ? gpu_uuids[i] : "";
    if (bits & 4) {
      snprintf(buf, sizeof buf, "%s has only %d °C margin left to slowdown (threshold %d °C)", uuid, ts[i].slowdown_margin_c, margin_threshold_c);
      margin.push_back(buf);
    }
    if (bits & 1) {
      snprintf(buf, sizeof buf, "%s current temperature is %u °C exceeding the threshold %u °C", uuid, ts[i].current_gpu_core_c, ts[i].threshold_gpu_max_c);
      gpu.push_back(buf);
    }
    if (bits & 2) {
      snprintf(buf, sizeof buf, "%s HBM temperature is %u °C exceeding the threshold %u °C", uuid, ts[i].current_hbm_c, ts[i].threshold_mem_max_c);
      hbm.push_back(buf);
    }
  }
  auto join = [](const std::vector<std::string>& v) { std::string j; for (size_t i = 0; i < v.size(); ++i) { if (i) j += ", "; j += v[i]; } return j; };
  std::string o;
  int32_t h = 1;
  if (!margin.empty()) o = "margin threshold exceeded: " + join(margin);
  else if (!gpu.empty()) o = "GPU temperature anomalies detected: " + join(gpu);
  else if (!hbm.empty()) o = "HBM temperature anomalies detected: " + join(hbm);
  else { h = 0; o = "all " + std::to_string(n) + " GPU(s) were checked, no temperature issue found"; }
  if (health) *health = h;
  if ((int32_t)o.size() + 1 > cap) return -1;
  memcpy(out, o.c_str(), o.size() + 1);
  return (int32_t)o.size();
}

namespace gpud {
// updateCurrentState (xid/component.go:581-611, sxid/component.go:478-507) over the SQLite stores: reboot events of the "os" bucket
// (pkg/host/event.go:15-17,134-158: name "reboot", newest first) and the component's own events since now - lookback, the latter cut at
// the newest SetHealthy, merged newest first, folded by evolveHealthyState.
int32_t read_bucket(gpud_store* st, const char* table, int64_t since, bool only_reboots, std::vector<Event>* out) {
  std::vector<gpud_event_row> rows(256);
  std::vector<char> text(1 << 16);
  int32_t n = 0;
  for (;;) {
    const int32_t rc = gpud_store_get_events(st, table, since, rows.data(), (int32_t)rows.size(), text.data(), (int32_t)text.size(), &n);
    if (rc == GPUD_OK) break;
    if (rc != GPUD_E_CAPACITY || rows.size() > (1u << 22)) return rc;
    rows.resize(rows.size() * 4);
    text.resize(text.size() * 4);
  }
  for (int32_t i = 0; i < n; ++i) {
    if (only_reboots && strcmp(rows[(size_t)i].name, "reboot") != 0) continue;
    Event e;
    e.time_unix = rows[(size_t)i].unix_s;
    e.name = rows[(size_t)i].name;
    e.type = rows[(size_t)i].type;
    e.message = std::string(text.data() + rows[(size_t)i].message_off, (size_t)rows[(size_t)i].message_len);
    // ExtraInfo of an error event: "data" and "device_uuid" (xid/component.go:536-540); read with the same JSON reader as Find
    const std::string extra(text.data() + rows[(size_t)i].extra_off, (size_t)rows[(size_t)i].extra_len);
    const char* p = extra.c_str();
    jsonmin::ws(p);
    if (*p == '{') {
      ++p;
      for (;;) {
        std::string k, v;
        jsonmin::ws(p);
        if (*p == '}' || !jsonmin::string(p, &k)) break;
        jsonmin::ws(p);
        if (*p++ != ':') break;
        jsonmin::ws(p);
        if (!strncmp(p, "null", 4)) p += 4;
        else if (!jsonmin::string(p, &v)) break;
        e.extra_info[k] = v;
        jsonmin::ws(p);
        if (*p == ',') { ++p; continue; }
        break;
      }
    }
    out->push_back(e);
  }
  return GPUD_OK;
}

int32_t state_from_store(gpud_store* st, const char* table, const char* os_table, int64_t now_unix, int64_t lookback_seconds, bool sxid, int reboot_threshold,
                                const DeviceMap& devices, int32_t* health, int32_t* action, char* reason, int32_t cap) {
  if (!st || !table || !health || !reason || cap <= 0) return GPUD_E_INVALID;
  const int64_t since = now_unix - lookback_seconds;
  std::vector<Event> reboots, local;
  if (os_table && *os_table) {
    const int32_t rc = read_bucket(st, os_table, since, true, &reboots);
    if (rc) return rc;
  }
  const int32_t rc = read_bucket(st, table, since, false, &local);
  if (rc) return rc;
  const std::vector<Event> all = merge_events(reboots, trim_events_after_set_healthy(local));
  const StoredEvolveResult r = sxid ? evolve_stored_sxid_events(all) : evolve_stored_events(all, devices, reboot_threshold);
  *health = (int32_t)r.health;
  if (action) *action = r.has_actions && !r.actions.empty() ? r.actions[0] : 0;
  if ((int32_t)r.reason.size() + 1 > cap) return GPUD_E_CAPACITY;
  memcpy(reason, r.reason.c_str(), r.reason.size() + 1);
  return GPUD_OK;
}
}  // namespace gpud

// ClockEvents.HWSlowdownEvent (hw-slowdown/clock_events.go:87-102): the Message of the "hw_slowdown" event for one reading -- the
// sorted hardware-slowdown reasons, each prefixed "<uuid>: " (:158-161), joined by ", ".  Returns the length; 0 = no event.
extern "C" int32_t gpud_hw_slowdown_event_message(uint64_t bitmask, const char* gpu_uuid, char* out, int32_t cap) {
  if (!out || cap <= 0) return -1;
  std::vector<std::string> hw;
  for (const ClockReason& r : kClockReasons)
    if ((bitmask & r.flag) && r.hw) hw.push_back(r.text);
  std::sort(hw.begin(), hw.end());
  std::string j;
  for (size_t i = 0; i < hw.size(); ++i) {
    if (i) j += ", ";
    j += gpu_uuid ? gpu_uuid : "";
    j += ": ";
    j += hw[i];
  }
  if ((int32_t)j.size() + 1 > cap) return -1;
  memcpy(out, j.c_str(), j.size() + 1);
  return (int32_t)j.size();
}

// The evaluation half of the hw-slowdown Check (hw-slowdown/component.go:352-407) over the event times read back from the bucket.
extern "C" int32_t gpud_hw_slowdown_check(const int64_t* event_unix, int32_t n, int64_t now_unix, int64_t window_seconds, double threshold_per_minute,
                                          int32_t* health, double* freq_per_minute, int32_t* hardware_inspection, char* reason, int32_t reason_cap) {
  if (n < 0 || (n && !event_unix) || !health) return GPUD_E_INVALID;
  const gpud::SlowdownVerdict v = gpud::evaluate_hw_slowdown(std::vector<int64_t>(event_unix, event_unix + n), now_unix, window_seconds, threshold_per_minute);
  *health = (int32_t)v.health;
  if (freq_per_minute) *freq_per_minute = v.freq_per_min;
  if (hardware_inspection) *hardware_inspection = v.inspect ? 1 : 0;
  if (reason) {
    if ((int32_t)v.reason.size() + 1 > reason_cap) return GPUD_E_CAPACITY;
    memcpy(reason, v.reason.c_str(), v.reason.size() + 1);
  }
  return GPUD_OK;
}

// ---- the kmsg watcher's duplicate drop (pkg/kmsg/watcher.go:281-286) over the units of a scanned buffer ----------------------------
// readFollow / ReadAll hand a parsed message to the matchers only the first time its (minute, message) key shows up within the
// cache TTL (deduper.go:63-125).  dropped[u] = 1 for every unit the watcher would have skipped; the cache lives in `d` across calls.
extern "C" int32_t gpud_kmsg_dedup_units(void* d, const uint8_t* buf, int64_t len, int32_t mode, int64_t boot_unix, int64_t lines_unix, int64_t now_unix,
                                         uint8_t* dropped, int64_t n_units, int64_t* n_dropped) {
  if (!d || len < 0 || (len && !buf) || !dropped || n_units < 0) return GPUD_E_INVALID;
  gpud::Deduper* dd = static_cast<gpud::Deduper*>(d);
  const bool raw = (mode & 0xff) == GPUD_SCAN_RAW_KMSG;
  int64_t u = 0, start = 0, nd = 0;
  for (int64_t i = 0; i <= len && u < n_units; ++i) {
    const bool sep = i == len || (buf[i] == '\n' && (!raw || i + 1 >= len || buf[i + 1] != ' '));
    if (!sep) continue;
    const std::string unit(reinterpret_cast<const char*>(buf) + start, (size_t)(i - start));
    dropped[u] = 0;
    if (!unit.empty()) {                                       // `if len(line) == 0 { continue }` (watcher.go:271-273)
      if (raw) {
        gpud::KmsgMessage m;
        std::string err;
        if (gpud::parse_kmsg_line(unit, &m, &err)) dropped[u] = dd->add(now_unix, boot_unix + m.usec_since_boot / 1000000, m.message) > 1;
      } else {
        dropped[u] = dd->add(now_unix, lines_unix, unit) > 1;
      }
    }
    nd += dropped[u];
    ++u;
    start = i + 1;
  }
  for (; u < n_units; ++u) dropped[u] = 0;
  if (n_dropped) *n_dropped = nd;
  return GPUD_OK;
}

// ---- flat C entry points so the host mirror is testable through ctypes (not part of gpud_b200.h) ----
extern "C" {

int32_t gpudh_parse_kmsg_line(const char* line, int32_t* prio, int64_t* seq, int64_t* usec, char* msg, int32_t cap) {
  gpud::KmsgMessage m;
  std::string err;
  if (!gpud::parse_kmsg_line(line, &m, &err)) return -1;
  *prio = m.priority; *seq = m.sequence; *usec = m.usec_since_boot;
  snprintf(msg, (size_t)cap, "%s", m.message.c_str());
  return 0;
}
int32_t gpudh_dedup_key(int64_t unix_s, const char* msg, char* out, int32_t cap) {
  snprintf(out, (size_t)cap, "%s", gpud::dedup_key(unix_s, msg).c_str());
  return 0;
}
void* gpudh_deduper_new(int64_t ttl) { return new gpud::Deduper(ttl); }
void* gpudh_deduper_new2(int64_t ttl, int32_t truncate_seconds) { return new gpud::Deduper(ttl, truncate_seconds > 0 ? truncate_seconds : 60); }   // WithCacheKeyTruncateSeconds
int32_t gpudh_deduper_add(void* d, int64_t now, int64_t t, const char* msg) { return static_cast<gpud::Deduper*>(d)->add(now, t, msg); }
void gpudh_deduper_free(void* d) { delete static_cast<gpud::Deduper*>(d); }
void* gpud_kmsg_deduper_create(int64_t ttl, int32_t truncate_seconds) { return new gpud::Deduper(ttl > 0 ? ttl : 15 * 60, truncate_seconds > 0 ? truncate_seconds : 60); }
void gpud_kmsg_deduper_destroy(void* d) { delete static_cast<gpud::Deduper*>(d); }

// events: n records of {kind 0 xid / 1 reboot / 2 SetHealthy(ignored here), event_type, xid, n_actions(-1 nil), actions[4]} newest first
typedef struct { int32_t kind, event_type; uint64_t xid; int32_t n_actions; int32_t actions[4]; } gpudh_event;
static int32_t evolve_flat(const gpudh_event* ev, int32_t n, int32_t reboot_threshold, const char* err_name, int32_t* health, int32_t* action, uint64_t* xid);
int32_t gpudh_evolve(const gpudh_event* ev, int32_t n, int32_t reboot_threshold, int32_t* health, int32_t* action, uint64_t* xid) {
  return evolve_flat(ev, n, reboot_threshold, "error_xid", health, action, xid);
}
int32_t gpudh_evolve_sxid(const gpudh_event* ev, int32_t n, int32_t* health, int32_t* action, uint64_t* sxid) {
  return evolve_flat(ev, n, 2 /* sxid/health_state.go:36 */, "error_sxid", health, action, sxid);
}
static int32_t evolve_flat(const gpudh_event* ev, int32_t n, int32_t reboot_threshold, const char* err_name, int32_t* health, int32_t* action, uint64_t* xid) {
  std::vector<gpud::XidEventView> v;
  for (int i = 0; i < n; ++i) {
    gpud::XidEventView e;
    e.name = ev[i].kind == 0 ? err_name : (ev[i].kind == 1 ? "reboot" : "SetHealthy");
    e.type = gpud::event_type_string(ev[i].event_type);
    e.xid = ev[i].xid;
    e.has_actions = ev[i].n_actions >= 0;
    for (int k = 0; k < ev[i].n_actions && k < 4; ++k) e.actions.push_back(ev[i].actions[k]);
    v.push_back(e);
  }
  const gpud::EvolveResult r = gpud::evolve_healthy_state(v, reboot_threshold, err_name);
  *health = (int32_t)r.health;
  *action = r.has_actions && !r.actions.empty() ? r.actions[0] : 0;
  *xid = r.has_xid ? r.xid : 0;
  return 0;
}

static gpud::DeviceMap parse_devices(const char* spec) {     // "uuid=busid;uuid=busid"
  gpud::DeviceMap m;
  std::string s = spec ? spec : "";
  size_t a = 0;
  while (a < s.size()) {
    size_t e = s.find(';', a);
    if (e == std::string::npos) e = s.size();
    const size_t q = s.find('=', a);
    if (q != std::string::npos && q < e) m[s.substr(a, q - a)] = s.substr(q + 1, e - q - 1);
    a = e + 1;
  }
  return m;
}
// health state of the xid / sxid component from the stores (see state_from_store above)
int32_t gpud_xid_state_from_store(gpud_store* st, const char* xid_table, const char* os_table, int64_t now_unix, int64_t lookback_seconds, int32_t reboot_threshold,
                                  const char* devices, int32_t* health, int32_t* action, char* reason, int32_t cap) {
  return gpud::state_from_store(st, xid_table, os_table, now_unix, lookback_seconds, false, reboot_threshold, parse_devices(devices), health, action, reason, cap);
}
int32_t gpud_sxid_state_from_store(gpud_store* st, const char* sxid_table, const char* os_table, int64_t now_unix, int64_t lookback_seconds, int32_t* health,
                                   int32_t* action, char* reason, int32_t cap) {
  return gpud::state_from_store(st, sxid_table, os_table, now_unix, lookback_seconds, true, 2, gpud::DeviceMap(), health, action, reason, cap);
}
// resolveXIDEvent for one stored event: 1 resolved (type_out / message_out filled), 0 left as is
int32_t gpudh_resolve_xid_event(const char* type_in, const char* raw, const char* device_uuid, const char* devices, char* type_out, int32_t tcap, char* msg_out,
                                int32_t mcap, int32_t* n_actions, int32_t* actions4) {
  std::string type = type_in ? type_in : "", msg;
  gpud::XidPayload p;
  if (!gpud::resolve_xid_event(&type, raw ? raw : "", device_uuid ? device_uuid : "", parse_devices(devices), &p, &msg)) return 0;
  snprintf(type_out, (size_t)tcap, "%s", type.c_str());
  snprintf(msg_out, (size_t)mcap, "%s", msg.c_str());
  *n_actions = p.has_actions ? (int32_t)p.actions.size() : -1;
  for (int i = 0; i < 4; ++i) actions4[i] = i < (int)p.actions.size() ? p.actions[(size_t)i] : 0;
  return 1;
}
// evolveHealthyState over stored events; events = records "name\x1ftype\x1fdevice_uuid\x1fdata" joined by \x1e, newest first
int32_t gpudh_evolve_stored(const char* events, const char* devices, int32_t reboot_threshold, int32_t* health, int32_t* action, char* reason, int32_t cap) {
  std::vector<gpud::Event> evs;
  std::string s = events ? events : "";
  size_t a = 0;
  while (a < s.size()) {
    size_t e = s.find('\x1e', a);
    if (e == std::string::npos) e = s.size();
    const std::string rec = s.substr(a, e - a);
    std::vector<std::string> f;
    size_t b = 0;
    for (;;) {
      const size_t q = rec.find('\x1f', b);
      f.push_back(rec.substr(b, q == std::string::npos ? std::string::npos : q - b));
      if (q == std::string::npos) break;
      b = q + 1;
    }
    f.resize(4);
    gpud::Event ev;
    ev.name = f[0]; ev.type = f[1];
    if (ev.name == "error_xid") { ev.extra_info["device_uuid"] = f[2]; ev.extra_info["data"] = f[3]; }
    evs.push_back(ev);
    a = e + 1;
  }
  const gpud::StoredEvolveResult r = gpud::evolve_stored_events(evs, parse_devices(devices), reboot_threshold);
  *health = (int32_t)r.health;
  *action = r.has_actions && !r.actions.empty() ? r.actions[0] : 0;
  snprintf(reason, (size_t)cap, "%s", r.reason.c_str());
  return 0;
}

int32_t gpudh_evolve_stored_sxid(const char* events, int32_t* health, int32_t* action, char* reason, int32_t cap) {
  std::vector<gpud::Event> evs;
  std::string s = events ? events : "";
  size_t a = 0;
  while (a < s.size()) {
    size_t e = s.find('\x1e', a);
    if (e == std::string::npos) e = s.size();
    const std::string rec = s.substr(a, e - a);
    std::vector<std::string> f;
    size_t b = 0;
    for (;;) {
      const size_t q = rec.find('\x1f', b);
      f.push_back(rec.substr(b, q == std::string::npos ? std::string::npos : q - b));
      if (q == std::string::npos) break;
      b = q + 1;
    }
    f.resize(4);
    gpud::Event ev;
    ev.name = f[0]; ev.type = f[1];
    if (ev.name == "error_sxid") { ev.extra_info["device_uuid"] = f[2]; ev.extra_info["data"] = f[3]; }
    evs.push_back(ev);
    a = e + 1;
  }
  const gpud::StoredEvolveResult r = gpud::evolve_stored_sxid_events(evs);
  *health = (int32_t)r.health;
  *action = r.has_actions && !r.actions.empty() ? r.actions[0] : 0;
  snprintf(reason, (size_t)cap, "%s", r.reason.c_str());
  return 0;
}

// mergeEvents / trimEventsAfterSetHealthy (xid/component.go:614-642) on bare (time, name) lists; names joined by '\n'
int32_t gpudh_merge_times(const int64_t* a, int32_t na, const int64_t* b, int32_t nb, int64_t* out) {
  std::vector<gpud::Event> ea((size_t)na), eb((size_t)nb);
  for (int32_t i = 0; i < na; ++i) ea[(size_t)i].time_unix = a[i];
  for (int32_t i = 0; i < nb; ++i) eb[(size_t)i].time_unix = b[i];
  const std::vector<gpud::Event> m = gpud::merge_events(ea, eb);
  for (size_t i = 0; i < m.size(); ++i) out[i] = m[i].time_unix;
  return (int32_t)m.size();
}
int32_t gpudh_trim_count(const char* names_newest_first) {
  std::vector<gpud::Event> ev;
  std::string s = names_newest_first ? names_newest_first : "";
  size_t a = 0;
  while (a <= s.size() && !s.empty()) {
    size_t e = s.find('\n', a);
    if (e == std::string::npos) e = s.size();
    gpud::Event x;
    x.name = s.substr(a, e - a);
    ev.push_back(x);
    if (e == s.size()) break;
    a = e + 1;
  }
  return (int32_t)gpud::trim_events_after_set_healthy(ev).size();
}

void* gpudh_xid_component_new(gpud_ctx* ctx, int32_t dev, int32_t row_remap, int32_t reboot_threshold) {
  return new gpud::XidComponent(ctx, dev, row_remap != 0, reboot_threshold);
}
void gpudh_xid_component_free(void* c) { delete static_cast<gpud::XidComponent*>(c); }
void gpudh_xid_component_set_source(void* c, const char* buf, int64_t len, int32_t raw, int64_t boot_unix) {
  static_cast<gpud::XidComponent*>(c)->SetKmsgSource(std::string(buf, (size_t)len), raw != 0, boot_unix);
}
// Check(): returns the number of found errors, health (0/1/2) and the summary string
int32_t gpudh_xid_component_check(void* c, int32_t* health, char* summary, int32_t cap, int32_t ingest, int64_t now_unix) {
  auto* x = static_cast<gpud::XidComponent*>(c);
  gpud::CheckResult cr = x->Check();
  *health = (int32_t)cr.health;
  snprintf(summary, (size_t)cap, "%s", cr.summary.c_str());
  if (ingest) x->IngestHits(cr.found, now_unix);
  return (int32_t)cr.found.size();
}
int32_t gpudh_xid_component_state_json(void* c, char* out, int32_t cap) {
  auto s = static_cast<gpud::XidComponent*>(c)->LastHealthStates();
  snprintf(out, (size_t)cap, "%s", s[0].to_json().c_str());
  return 0;
}
int32_t gpudh_xid_component_reboot(void* c, int64_t t) { static_cast<gpud::XidComponent*>(c)->AddRebootEvent(t); static_cast<gpud::XidComponent*>(c)->Start(); return 0; }
int32_t gpudh_xid_component_set_healthy(void* c, int64_t t) { return static_cast<gpud::XidComponent*>(c)->SetHealthy(t); }
int32_t gpudh_xid_component_n_events(void* c) { return (int32_t)static_cast<gpud::XidComponent*>(c)->Events(INT64_MIN).size(); }
int32_t gpudh_xid_component_n_events_since(void* c, int64_t since) { return (int32_t)static_cast<gpud::XidComponent*>(c)->Events(since).size(); }
const char* gpudh_xid_component_name(void) { return gpud::XidComponent::kName; }
int32_t gpudh_hw_slowdown(const int64_t* ev, int32_t n, int64_t now, int64_t window_s, double thr, double* freq, int32_t* distinct) {
  const gpud::SlowdownVerdict v = gpud::evaluate_hw_slowdown(std::vector<int64_t>(ev, ev + n), now, window_s, thr);
  *freq = v.freq_per_min; *distinct = v.distinct_minutes;
  return (int32_t)v.health;
}
void gpudh_go_duration(int64_t sec, char* out, int32_t cap) { snprintf(out, (size_t)cap, "%s", gpud::go_duration_seconds(sec).c_str()); }
int32_t gpudh_hw_slowdown_reason(const int64_t* ev, int32_t n, int64_t now, int64_t window_s, double thr, char* out, int32_t cap) {
  const gpud::SlowdownVerdict v = gpud::evaluate_hw_slowdown(std::vector<int64_t>(ev, ev + n), now, window_s, thr);
  snprintf(out, (size_t)cap, "%s", v.reason.c_str());
  return (int32_t)v.inspect;
}
}  // extern "C" (flat test entries)

// public form of the temperature rules over a poller reading
extern "C" int32_t gpud_temperature_check(const gpud_temperature* t, int32_t margin_threshold_c, int32_t* bits) {
  if (!t || !bits) return GPUD_E_INVALID;
  gpud::TemperatureReading r;
  r.current_gpu_core = t->current_gpu_core_c; r.threshold_gpu_max = t->threshold_gpu_max_c; r.current_hbm = t->current_hbm_c;
  r.threshold_mem_max = t->threshold_mem_max_c; r.hbm_supported = t->hbm_supported != 0; r.threshold_slowdown = t->threshold_slowdown_c;
  r.slowdown_margin = t->slowdown_margin_c; r.margin_supported = t->margin_supported != 0;
  *bits = gpud::evaluate_temperature(r, margin_threshold_c);
  return GPUD_OK;
}

extern "C" {
int32_t gpudh_temperature(uint32_t cur, uint32_t gmax, uint32_t hbm, uint32_t mmax, int32_t hbm_supported, uint32_t slowdown, int32_t margin, int32_t margin_supported,
                          int32_t mthr) {
  gpud::TemperatureReading t;
  t.current_gpu_core = cur; t.threshold_gpu_max = gmax; t.current_hbm = hbm; t.threshold_mem_max = mmax; t.hbm_supported = hbm_supported != 0;
  t.threshold_slowdown = slowdown; t.slowdown_margin = margin; t.margin_supported = margin_supported != 0;
  return gpud::evaluate_temperature(t, mthr);
}
}
