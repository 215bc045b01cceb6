#include "../../lpref_ros.hpp"
