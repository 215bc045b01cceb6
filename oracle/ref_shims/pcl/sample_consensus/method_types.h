#include "../../lpref_pcl.hpp"
