/* stub of <libconfig.h++> for `make refcheck`: the reference headers only name these types */
#pragma once
namespace libconfig {
class Setting;
class Config;
}
