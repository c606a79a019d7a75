// TEST INFRASTRUCTURE: see core.hpp (everything the reference TUs need lives there).
#include "core.hpp"
