#pragma once
#include <tf/tf.h>
namespace tf { struct TransformBroadcaster { void sendTransform(const StampedTransform &) {} }; }
