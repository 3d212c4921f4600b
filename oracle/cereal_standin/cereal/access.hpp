// Header-only stand-in for the (absent) cereal library: just enough declarations for the
// reference's hot-path sources to compile when no archive is ever instantiated.
// Test infrastructure only (used to build oracle/_ref); not part of the product.
#pragma once
namespace cereal {
class access {};
namespace specialization { struct non_member_load_save {}; }
template <class Archive, class T, class S> struct specialize {};
}
#ifndef CEREAL_REGISTER_TYPE
#define CEREAL_REGISTER_TYPE(T)
#endif
#ifndef CEREAL_REGISTER_POLYMORPHIC_RELATION
#define CEREAL_REGISTER_POLYMORPHIC_RELATION(B, D)
#endif
