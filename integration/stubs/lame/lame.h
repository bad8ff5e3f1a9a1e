/* stub: just enough of <lame/lame.h> for the reference's rtl_airband.h to compile in `make refcheck` (layout checks only) */
#pragma once
typedef struct lame_global_struct lame_global_flags;
typedef lame_global_flags* lame_t;
