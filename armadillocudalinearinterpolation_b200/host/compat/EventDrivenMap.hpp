// Source-compatibility header: code written against the reference's EventDrivenMap
// (e.g. its Driver.cu, which includes "EventDrivenMap.hpp" and "parameters.hpp") builds
// against the B200 map without edits.
#ifndef EVENTDRIVEMAPHEADERDEF
#define EVENTDRIVEMAPHEADERDEF
#include "EventDrivenMapB200.hpp"
typedef EventDrivenMapB200 EventDrivenMap;
#endif
