/*
 * include/user.h -- the operator knobs, same macro names as the reference's
 * sort-merge-join/user.h:1-13.  NR_DPUS / NR_TASKLETS (user.h:3-4) are replaced
 * by NR_GPUS.  These are compile-time DEFAULTS only: smj_config_default() copies
 * them into a run-time smj_config_t, and host/app.c lets the command line or the
 * environment (SMJ_NR_GPUS, SMJ_SELECT_VAL1, ...) override them without a rebuild.
 */
#ifndef SMJ_USER_H
#define SMJ_USER_H

// #define DEBUG

#ifndef NR_GPUS
#define NR_GPUS 1
#endif

#ifndef SELECT_COL1
#define SELECT_COL1 0
#endif
#ifndef SELECT_VAL1
#define SELECT_VAL1 5000
#endif

#ifndef SELECT_COL2
#define SELECT_COL2 0
#endif
#ifndef SELECT_VAL2
#define SELECT_VAL2 5000
#endif

#ifndef JOIN_KEY1
#define JOIN_KEY1 0
#endif
#ifndef JOIN_KEY2
#define JOIN_KEY2 0
#endif

#endif /* SMJ_USER_H */
