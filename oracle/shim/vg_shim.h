// TEST INFRASTRUCTURE ONLY (oracle build). Not part of the product.
//
// Stand-in for the protoc-generated vg.pb.h and for stream.hpp so that the
// reference's hot-path sources compile *in place* from /root/reference without
// protobuf.  It is force-included (-include) ahead of every reference TU and
// pre-defines the two include guards (vg.pb.h:4, stream.hpp:1) so the
// reference's own copies of those headers expand to nothing.
//
// Only the accessors the reference actually calls are provided (see
// GraphAligner.h:782-847, Aligner.cpp:24-91, BigraphToDigraph.cpp:27-57).
#ifndef GA_ORACLE_VG_SHIM_H
#define GA_ORACLE_VG_SHIM_H
#define PROTOBUF_vg_2eproto__INCLUDED
#define STREAM_H

#include <cassert>
#include <cstdint>
#include <deque>
#include <fstream>
#include <functional>
#include <iostream>
#include <istream>
#include <memory>
#include <ostream>
#include <string>
#include <vector>

namespace vg {

// deep-copying owning pointer: set_allocated_* hands over a raw pointer which
// the reference keeps writing through afterwards (GraphAligner.h:788-789,804-810)
template <typename T>
class owned {
public:
	owned() : p(nullptr) {}
	owned(const owned& o) : p(o.p ? new T(*o.p) : nullptr) {}
	owned& operator=(const owned& o) { if (this != &o) { T* n = o.p ? new T(*o.p) : nullptr; delete p; p = n; } return *this; }
	~owned() { delete p; }
	void reset(T* n) { if (n != p) delete p; p = n; }
	T* get() { if (!p) p = new T(); return p; }
	const T& cget() const { static const T empty{}; return p ? *p : empty; }
private:
	T* p;
};

class Position {
public:
	Position() : node_id_(0), offset_(0), is_reverse_(false) {}
	int64_t node_id() const { return node_id_; }
	void set_node_id(int64_t v) { node_id_ = v; }
	int64_t offset() const { return offset_; }
	void set_offset(int64_t v) { offset_ = v; }
	bool is_reverse() const { return is_reverse_; }
	void set_is_reverse(bool v) { is_reverse_ = v; }
	const std::string& name() const { return name_; }
	void set_name(const std::string& v) { name_ = v; }
private:
	int64_t node_id_; int64_t offset_; bool is_reverse_; std::string name_;
};

class Edit {
public:
	Edit() : from_length_(0), to_length_(0) {}
	int32_t from_length() const { return from_length_; }
	void set_from_length(int32_t v) { from_length_ = v; }
	int32_t to_length() const { return to_length_; }
	void set_to_length(int32_t v) { to_length_ = v; }
	const std::string& sequence() const { return sequence_; }
	void set_sequence(const std::string& v) { sequence_ = v; }
private:
	int32_t from_length_; int32_t to_length_; std::string sequence_;
};

class Mapping {
public:
	Mapping() : rank_(0) {}
	const Position& position() const { return position_.cget(); }
	Position* mutable_position() { return position_.get(); }
	void set_allocated_position(Position* p) { position_.reset(p); }
	Edit* add_edit() { edits_.emplace_back(); return &edits_.back(); }
	int edit_size() const { return (int)edits_.size(); }
	const Edit& edit(int i) const { return edits_[i]; }
	Edit* mutable_edit(int i) { return &edits_[i]; }
	int64_t rank() const { return rank_; }
	void set_rank(int64_t v) { rank_ = v; }
private:
	owned<Position> position_; std::deque<Edit> edits_; int64_t rank_;
};

class Path {
public:
	Path() : is_circular_(false), length_(0) {}
	const std::string& name() const { return name_; }
	void set_name(const std::string& v) { name_ = v; }
	Mapping* add_mapping() { mappings_.emplace_back(); return &mappings_.back(); }
	int mapping_size() const { return (int)mappings_.size(); }
	const Mapping& mapping(int i) const { return mappings_[i]; }
	Mapping* mutable_mapping(int i) { return &mappings_[i]; }
private:
	std::string name_; std::deque<Mapping> mappings_; bool is_circular_; int64_t length_;
};

class Alignment {
public:
	Alignment() : score_(0), query_position_(0), mapping_quality_(0) {}
	const std::string& sequence() const { return sequence_; }
	void set_sequence(const std::string& v) { sequence_ = v; }
	const std::string& name() const { return name_; }
	void set_name(const std::string& v) { name_ = v; }
	const std::string& quality() const { return quality_; }
	void set_quality(const std::string& v) { quality_ = v; }
	int32_t score() const { return score_; }
	void set_score(int32_t v) { score_ = v; }
	int32_t query_position() const { return query_position_; }
	void set_query_position(int32_t v) { query_position_ = v; }
	const Path& path() const { return path_.cget(); }
	Path* mutable_path() { return path_.get(); }
	void set_allocated_path(Path* p) { path_.reset(p); }
private:
	std::string sequence_, name_, quality_; int32_t score_, query_position_, mapping_quality_; owned<Path> path_;
};

class Node {
public:
	Node() : id_(0) {}
	const std::string& sequence() const { return sequence_; }
	void set_sequence(const std::string& v) { sequence_ = v; }
	const std::string& name() const { return name_; }
	void set_name(const std::string& v) { name_ = v; }
	int64_t id() const { return id_; }
	void set_id(int64_t v) { id_ = v; }
private:
	std::string sequence_, name_; int64_t id_;
};

class Edge {
public:
	Edge() : from_(0), to_(0), from_start_(false), to_end_(false), overlap_(0) {}
	int64_t from() const { return from_; }
	void set_from(int64_t v) { from_ = v; }
	int64_t to() const { return to_; }
	void set_to(int64_t v) { to_ = v; }
	bool from_start() const { return from_start_; }
	void set_from_start(bool v) { from_start_ = v; }
	bool to_end() const { return to_end_; }
	void set_to_end(bool v) { to_end_ = v; }
	int32_t overlap() const { return overlap_; }
	void set_overlap(int32_t v) { overlap_ = v; }
private:
	int64_t from_, to_; bool from_start_, to_end_; int32_t overlap_;
};

class Graph {
public:
	Node* add_node() { nodes_.emplace_back(); return &nodes_.back(); }
	int node_size() const { return (int)nodes_.size(); }
	const Node& node(int i) const { return nodes_[i]; }
	Edge* add_edge() { edges_.emplace_back(); return &edges_.back(); }
	int edge_size() const { return (int)edges_.size(); }
	const Edge& edge(int i) const { return edges_[i]; }
private:
	std::deque<Node> nodes_; std::deque<Edge> edges_;
};

} // namespace vg

// stream.hpp stand-in: the library-only oracle never touches files; the
// functions exist so CommonUtils.cpp links, and refuse to run.
namespace stream {
template <typename T>
bool for_each(std::istream&, std::function<void(T&)>&) {
	std::cerr << "oracle shim: stream::for_each is not available in the library-only oracle" << std::endl;
	std::abort();
}
template <typename T>
bool write_buffered(std::ostream&, std::vector<T>&, uint64_t) {
	std::cerr << "oracle shim: stream::write_buffered is not available in the library-only oracle" << std::endl;
	std::abort();
}
}

#endif
