// host_component.h — C++ host side above the C ABI: mirrors the reference's operator interface for this path
// (the reference is compiled Go and no Go toolchain exists in the build image, so the mirror is C++).
//
//   components.Component / CheckResult      components/types.go:20-100
//   api/v1 HealthState / Event / enums      api/v1/types.go:17-259
//   kmsg.parseLine / deduper                pkg/kmsg/watcher.go:292-332, pkg/kmsg/deduper.go:63-125
//   eventstore.Bucket (Insert/Find/Get)     pkg/eventstore/types.go:23-70, database.go:248-365,459-469
//   xid component Check / evolveHealthyState  components/accelerator/nvidia/xid/component.go:255-311, health_state.go:57-128
#pragma once
#include <stdint.h>

#include <map>
#include <memory>
#include <mutex>
#include <string>
#include <vector>

#include "../../include/gpud_b200.h"

namespace gpud {

// ---- api/v1 (types.go:17-259): wire strings are the reference's ----
enum class Health { Healthy = 0, Degraded = 1, Unhealthy = 2, Initializing = 3 };
const char* health_string(Health h);            // "Healthy" | "Degraded" | "Unhealthy" | "Initializing"
const char* event_type_string(int32_t t);       // GPUD_EVENT_* -> "Unknown" | "Info" | "Warning" | "Critical" | "Fatal"
const char* repair_action_string(int32_t a);    // GPUD_ACT_*   -> "REBOOT_SYSTEM" ...

void jstr(std::string& out, const std::string& s);   // append s as a JSON string the way encoding/json escapes it
struct SuggestedActions { std::vector<int32_t> repair_actions; };
struct HealthState {
  int64_t time_unix = 0;                          // 0 = the zero time (JSON null)
  std::string name, component, reason, error;
  Health health = Health::Healthy;
  bool has_actions = false;
  SuggestedActions actions;
  std::string to_json() const;
};
struct Event {                                   // eventstore.Event (pkg/eventstore/types.go:23-37)
  int64_t time_unix = 0;
  std::string name, type, message;
  std::map<std::string, std::string> extra_info;
};

// ---- pkg/kmsg ----
struct KmsgMessage { int priority = 0; int64_t sequence = 0; int64_t usec_since_boot = 0; std::string message; };
bool parse_kmsg_line(const std::string& line, KmsgMessage* out, std::string* err);   // watcher.go:292-332
std::string dedup_key(int64_t unix_seconds, const std::string& message, int truncate_seconds = 60);   // deduper.go:63-74
class Deduper {                                  // deduper.go:77-125 (go-cache with TTL; time is injected)
 public:
  explicit Deduper(int64_t ttl_seconds = 15 * 60, int truncate_seconds = 60) : ttl_(ttl_seconds), trunc_(truncate_seconds) {}
  int add(int64_t now_unix, int64_t msg_unix, const std::string& message);   // occurrence count: 1 = first
 private:
  struct Entry { int count; int64_t expires; };
  int64_t ttl_;
  int trunc_;
  std::map<std::string, Entry> cache_;
};

// ---- eventstore.Bucket semantics, in memory (the SQLite file stays gpud's; this is the sink contract) ----
class EventBucket {
 public:
  bool insert(const Event& ev);                                   // database.go:248-276
  const Event* find(const Event& ev) const;                       // database.go:278-324 + compareEvent :459-469
  std::vector<Event> get(int64_t since_unix) const;               // database.go:327-365: time DESC
  int purge(int64_t before_unix);                                 // database.go:149-166
  size_t size() const { return events_.size(); }
 private:
  std::vector<Event> events_;
};

// ---- xid health evolution (health_state.go:57-128; component.go:614-642) ----
// xidErrorEventDetail (xid/health_state.go:284-315) as read back from a stored event's "data" payload
struct XidPayload {
  uint64_t xid = 0;
  int32_t sub_code = 0;
  uint32_t error_status = 0;
  std::string device_uuid, sub_code_description, description;
  bool has_actions = false;                     // SuggestedActionsByGPUd != nil
  std::vector<int32_t> actions;
};
typedef std::map<std::string, std::string> DeviceMap;     // NVML UUID -> PCIBusID() ("0000:04:00.0")
// resolveXIDEvent + addEventDetails (xid/health_state.go:184-281): JSON payload (xid != 0) or the legacy decimal code; fills the
// fields the catalog Detail supplies, sets *type when it was "", renders *message with buildMessage.  false = the reference
// leaves the event unresolved (evolveHealthyState then skips it).
bool resolve_xid_event(std::string* type, const std::string& raw_data, const std::string& event_device_uuid, const DeviceMap& devices, XidPayload* out,
                       std::string* message);
std::string convert_bus_id_to_uuid(const std::string& bus_id, const DeviceMap& devices);      // health_state.go:171-182
std::string xid_payload_message(const XidPayload& p, const DeviceMap& devices);               // buildMessage, :130-169
struct StoredEvolveResult { Health health = Health::Healthy; bool has_actions = false; std::vector<int32_t> actions; std::string reason; };
// evolveHealthyState over stored events, newest first (health_state.go:57-128), reason included
StoredEvolveResult evolve_stored_events(const std::vector<Event>& events_newest_first, const DeviceMap& devices, int reboot_threshold);

// the sxid twins (sxid/health_state.go:38-142): the stored payload is the decimal code, resolved on read
struct SXidPayload { uint64_t sxid = 0; std::string device_uuid; bool has_actions = false; std::vector<int32_t> actions; };
bool resolve_sxid_event(std::string* type, const std::string& raw_data, const std::string& event_device_uuid, SXidPayload* out, std::string* message);
StoredEvolveResult evolve_stored_sxid_events(const std::vector<Event>& events_newest_first);

int32_t state_from_store(gpud_store* st, const char* table, const char* os_table, int64_t now_unix, int64_t lookback_seconds, bool sxid, int reboot_threshold,
                         const DeviceMap& devices, int32_t* health, int32_t* action, char* reason, int32_t cap);

struct XidEventView { std::string name; std::string type; uint64_t xid = 0; bool has_actions = false; std::vector<int32_t> actions; };
struct EvolveResult { Health health = Health::Healthy; bool has_actions = false; std::vector<int32_t> actions; bool has_xid = false; uint64_t xid = 0; int last_index = -1; /* index into the input of the event that is lastXidErr */ };
EvolveResult evolve_healthy_state(const std::vector<XidEventView>& events_newest_first, int reboot_threshold);
std::vector<Event> trim_events_after_set_healthy(const std::vector<Event>& events_newest_first);     // component.go:630-642
std::vector<Event> merge_events(const std::vector<Event>& a, const std::vector<Event>& b);           // component.go:614-628

// ---- windowed threshold rules with reference code ----
// hw-slowdown: distinct event-minutes in the evaluation window / window minutes >= threshold -> Unhealthy + HARDWARE_INSPECTION
// (components/accelerator/nvidia/hw-slowdown/component.go:352-407; defaults 10 min / 0.6 at :29-35)
struct SlowdownVerdict { Health health = Health::Healthy; double freq_per_min = 0.0; int distinct_minutes = 0; bool inspect = false; std::string reason; };
std::string go_duration_seconds(int64_t seconds);   // time.Duration.String() of a whole number of seconds
SlowdownVerdict evaluate_hw_slowdown(const std::vector<int64_t>& event_unix_seconds, int64_t now_unix, int64_t window_seconds, double threshold_per_min);
// temperature: current > max-operating, HBM > max-memory, margin <= configured threshold -> Degraded-class reasons
// (components/accelerator/nvidia/temperature/component.go:206-248); returns a bit mask 1 gpu, 2 hbm, 4 margin
struct TemperatureReading {                    // the fields of temperature.Temperature the rules read (temperature/temperature.go:17-60)
  uint32_t current_gpu_core = 0, current_hbm = 0;
  uint32_t threshold_slowdown = 0, threshold_mem_max = 0, threshold_gpu_max = 0;
  int32_t slowdown_margin = 0;
  bool hbm_supported = false, margin_supported = false;
};
int evaluate_temperature(const TemperatureReading& t, int32_t margin_threshold_c);

// ---- components.Component (types.go:20-66) ----
struct CheckResult {
  std::string component, summary, text;
  Health health = Health::Healthy;
  std::vector<HealthState> states;
  std::vector<gpud_xid_hit> found;               // FoundErrors of xid/sxid
};
class Component {
 public:
  virtual ~Component() {}
  virtual std::string Name() const = 0;
  virtual std::vector<std::string> Tags() const = 0;
  virtual bool IsSupported() const = 0;
  virtual int32_t Start() = 0;
  virtual CheckResult Check() = 0;
  virtual std::vector<HealthState> LastHealthStates() = 0;
  virtual std::vector<Event> Events(int64_t since_unix) = 0;
  virtual int32_t Close() = 0;
};

// The Xid component on the B200 path: Check() = scan the kmsg buffer on the GPU, drop 63/64 when row remapping is
// supported, Unhealthy iff any hit is Critical/Fatal (xid/component.go:255-311).  Streaming use feeds OnKmsg().
class XidComponent : public Component {
 public:
  static constexpr const char* kName = "accelerator-nvidia-error-xid";     // xid/component.go:35
  XidComponent(gpud_ctx* ctx, int32_t dev, bool row_remapping_supported, int reboot_threshold = 2);
  void SetKmsgSource(std::string buffer, bool raw_kmsg, int64_t boot_unix) { buf_ = std::move(buffer); raw_ = raw_kmsg; boot_unix_ = boot_unix; }
  void AddRebootEvent(int64_t unix_seconds);
  void SetDevices(DeviceMap devices) { std::lock_guard<std::mutex> g(mu_); devices_ = std::move(devices); }   // nvmlInstance.Devices() for the UUID in the reason
  std::string Name() const override { return kName; }
  std::vector<std::string> Tags() const override { return {"accelerator", "gpu", "nvidia", kName}; }
  bool IsSupported() const override { return ctx_ != nullptr; }
  int32_t Start() override;                       // non-blocking; state refresh only (component.go:141-170)
  CheckResult Check() override;
  std::vector<HealthState> LastHealthStates() override;
  std::vector<Event> Events(int64_t since_unix) override;
  int32_t Close() override { return 0; }
  int32_t SetHealthy(int64_t now_unix);           // set_healthy.go:14-35
  int32_t IngestHits(const std::vector<gpud_xid_hit>& hits, int64_t fallback_unix);   // start() body, component.go:468-577
 private:
  void update_state();                            // updateCurrentState, component.go:581-611
  gpud_ctx* ctx_;
  int32_t dev_;
  bool row_remap_;
  int reboot_threshold_;
  std::string buf_;
  bool raw_ = false;
  int64_t boot_unix_ = 0;
  std::mutex mu_;
  EventBucket bucket_, reboots_;
  DeviceMap devices_;
  HealthState cur_;
  bool checked_ = false;
};

}  // namespace gpud
