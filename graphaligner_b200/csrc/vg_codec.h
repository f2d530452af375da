// Protobuf-free codec for the vg "stream" container and the messages on the aligner's I/O surface.
// Container (reference stream.hpp:24-51 write, :69-111 read): gzip( varint64 count, {varint32 len, msg} x count )*.
// Messages / field numbers (reference vg.pb.h:149-960):
//   Graph{1 node,2 edge}  Node{1 sequence,2 name,3 id}  Edge{1 from,2 to,3 from_start,4 to_end,5 overlap}
//   Alignment{1 sequence,2 path,3 name,6 score,7 query_position}  Path{2 mapping}
//   Mapping{1 position,2 edit,5 rank}  Position{1 node_id,2 offset,4 is_reverse}  Edit{1 from_length,2 to_length,3 sequence}
// proto3: default-valued scalar fields are omitted on the wire.
#ifndef GA_VG_CODEC_H
#define GA_VG_CODEC_H
#include <string>
#include <vector>
#include "bigraph_to_digraph.h"
#include "ga_host.h"

namespace vgcodec
{
// every record of a stream file, undecoded
std::vector<std::string> ReadStreamFile(const std::string& filename);
void WriteStreamFile(const std::string& filename, const std::vector<std::string>& records);

void ReadGraphFile(const std::string& filename, std::vector<DirectedGraph::BiNode>& nodes, std::vector<DirectedGraph::BiEdge>& edges);
std::string EncodeGraph(const std::vector<DirectedGraph::BiNode>& nodes, const std::vector<DirectedGraph::BiEdge>& edges);

vg::Alignment DecodeAlignment(const std::string& msg);
std::string EncodeAlignment(const vg::Alignment& aln);
std::vector<vg::Alignment> ReadAlignmentFile(const std::string& filename);
void WriteAlignmentFile(const std::string& filename, const std::vector<vg::Alignment>& alns);
}

#endif
