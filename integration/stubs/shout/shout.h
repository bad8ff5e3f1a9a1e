/* stub of <shout/shout.h> for `make refcheck` */
#pragma once
typedef struct shout shout_t;
