/* gpud_b200_hooks.h — TEST HOOKS of libgpud_b200.so: flat C entry points over the pure host functions of the C++ mirror
 * (gpud_b200/csrc/host_component.h), exported so that the parity tests can drive each reference rule on its own through ctypes.
 * An integrator binds include/gpud_b200.h (the components go through gpud_component_*); nothing here is needed to use the
 * library, and nothing here touches a GPU except gpudh_xid_component_* (superseded by gpud_component_* and kept for the tests
 * that pin the in-memory bucket).  Every symbol the library exports is declared in one of the two headers (tests/test_abi_cpu.py). */
#ifndef GPUD_B200_HOOKS_H
#define GPUD_B200_HOOKS_H
#include "gpud_b200.h"
#ifdef __cplusplus
extern "C" {
#endif

/* pkg/kmsg: parseLine (watcher.go:292-332), the dedup key and cache (deduper.go:63-125) */
int32_t gpudh_parse_kmsg_line(const char* line, int32_t* prio, int64_t* seq, int64_t* usec, char* msg, int32_t cap);
int32_t gpudh_dedup_key(int64_t unix_s, const char* msg, char* out, int32_t cap);
void* gpudh_deduper_new(int64_t ttl);
void* gpudh_deduper_new2(int64_t ttl, int32_t truncate_seconds);
int32_t gpudh_deduper_add(void* d, int64_t now, int64_t t, const char* msg);
void gpudh_deduper_free(void* d);
/* evolveHealthyState (xid/health_state.go:57-128, sxid/health_state.go:38-111), resolveXIDEvent, mergeEvents / trimEventsAfterSetHealthy */
typedef struct { int32_t kind, event_type; uint64_t xid; int32_t n_actions; int32_t actions[4]; } gpudh_event;
int32_t gpudh_evolve(const gpudh_event* ev, int32_t n, int32_t reboot_threshold, int32_t* health, int32_t* action, uint64_t* xid);
int32_t gpudh_evolve_sxid(const gpudh_event* ev, int32_t n, int32_t* health, int32_t* action, uint64_t* sxid);
int32_t gpudh_resolve_xid_event(const char* type_in, const char* raw, const char* device_uuid, const char* devices, char* type_out, int32_t tcap, char* msg_out,
                                int32_t mcap, int32_t* n_actions, int32_t* actions4);
int32_t gpudh_evolve_stored(const char* events, const char* devices, int32_t reboot_threshold, int32_t* health, int32_t* action, char* reason, int32_t cap);
int32_t gpudh_evolve_stored_sxid(const char* events, int32_t* health, int32_t* action, char* reason, int32_t cap);
int32_t gpudh_merge_times(const int64_t* a, int32_t na, const int64_t* b, int32_t nb, int64_t* out);
int32_t gpudh_trim_count(const char* names_newest_first);
/* the xid component mirror with its in-memory bucket (tests of the bucket semantics; integrators use gpud_component_*) */
void* gpudh_xid_component_new(gpud_ctx* ctx, int32_t dev, int32_t row_remap, int32_t reboot_threshold);
void gpudh_xid_component_free(void* c);
void gpudh_xid_component_set_source(void* c, const char* buf, int64_t len, int32_t raw, int64_t boot_unix);
int32_t gpudh_xid_component_check(void* c, int32_t* health, char* summary, int32_t cap, int32_t ingest, int64_t now_unix);
int32_t gpudh_xid_component_state_json(void* c, char* out, int32_t cap);
int32_t gpudh_xid_component_reboot(void* c, int64_t t);
int32_t gpudh_xid_component_set_healthy(void* c, int64_t t);
int32_t gpudh_xid_component_n_events(void* c);
int32_t gpudh_xid_component_n_events_since(void* c, int64_t since_unix);
const char* gpudh_xid_component_name(void);
/* hw-slowdown window rule (hw-slowdown/component.go:352-407), Go's Duration.String, the temperature rule (temperature/component.go:206-248) */
int32_t gpudh_hw_slowdown(const int64_t* ev, int32_t n, int64_t now, int64_t window_s, double thr, double* freq, int32_t* distinct);
void gpudh_go_duration(int64_t sec, char* out, int32_t cap);
int32_t gpudh_hw_slowdown_reason(const int64_t* ev, int32_t n, int64_t now, int64_t window_s, double thr, char* out, int32_t cap);
int32_t gpudh_temperature(uint32_t cur, uint32_t gmax, uint32_t hbm, uint32_t mmax, int32_t hbm_supported, uint32_t slowdown, int32_t margin, int32_t margin_supported,
                          int32_t mthr);
/* pkg/nvidia/errors/error.go:33-127 (bit 1 not supported, 2 GPU lost, 4 reset required), the event store's SQL strings */
int32_t gpudh_nvml_error_class(int32_t ret, const char* error_string);
int32_t gpudh_store_event_sql(int32_t which, const char* table, char* out, int32_t cap);

#ifdef __cplusplus
}
#endif
#endif
