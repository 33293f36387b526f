/* fake_engine.h -- counters and fault injection of the C-ABI test double (fake_engine.cpp). */
#pragma once
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif
uint64_t fake_violations(void);        /* contract violations seen so far (overlapping calls, dead engine, bad index) */
uint64_t fake_engines_created(void);
uint64_t fake_engines_destroyed(void);
uint64_t fake_periods_processed(void); /* ca_process calls that ran */
uint64_t fake_ir_loads(void);
uint64_t fake_engines_on_device(int device); /* engines (and group members) created on that device ordinal */
uint64_t fake_groups_created(void);
void fake_set_process_delay_us(int us); /* every ca_process sleeps this long (widens the windows) */
void fake_set_create_delay_us(int us);  /* every ca_create sleeps this long (a build takes a while) */
void fake_fail_next_creates(int n);     /* the next n ca_create calls fail with CA_ERR_NOMEM */
#ifdef __cplusplus
}
#endif
