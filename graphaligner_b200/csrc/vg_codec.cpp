#include "vg_codec.h"
#include <zlib.h>
#include <algorithm>
#include <cstring>
#include <stdexcept>

namespace vgcodec
{

namespace
{

struct Reader
{
	const uint8_t* p;
	const uint8_t* end;
	bool done() const { return p >= end; }
	uint64_t varint()
	{
		uint64_t v = 0;
		int shift = 0;
		while (true)
		{
			if (p >= end) throw std::runtime_error("vg codec: truncated varint");
			uint8_t b = *p++;
			v |= (uint64_t)(b & 0x7f) << shift;
			if (!(b & 0x80)) return v;
			shift += 7;
			if (shift > 63) throw std::runtime_error("vg codec: varint too long");
		}
	}
	Reader sub()
	{
		uint64_t len = varint();
		if ((uint64_t)(end - p) < len) throw std::runtime_error("vg codec: truncated field");
		Reader r { p, p + len };
		p += len;
		return r;
	}
	std::string str()
	{
		Reader r = sub();
		return std::string((const char*)r.p, (size_t)(r.end - r.p));
	}
	void skip(int wireType)
	{
		if (wireType == 0) varint();
		else if (wireType == 2) sub();
		else if (wireType == 1 || wireType == 5)
		{
			const size_t n = wireType == 1 ? 8 : 4;
			if ((size_t)(end - p) < n) throw std::runtime_error("vg codec: truncated fixed-width field");
			p += n;
		}
		else throw std::runtime_error("vg codec: unsupported wire type");
	}
};

void putVarint(std::string& out, uint64_t v)
{
	while (v >= 0x80)
	{
		out.push_back((char)((v & 0x7f) | 0x80));
		v >>= 7;
	}
	out.push_back((char)v);
}

void putKey(std::string& out, int field, int wireType) { putVarint(out, (uint64_t)(field << 3 | wireType)); }

void putInt(std::string& out, int field, int64_t v)
{
	if (v == 0) return;
	putKey(out, field, 0);
	putVarint(out, (uint64_t)v);   // int32/int64: negative values are sign-extended to 10 bytes
}

void putBool(std::string& out, int field, bool v)
{
	if (!v) return;
	putKey(out, field, 0);
	out.push_back(1);
}

void putBytes(std::string& out, int field, const std::string& v, bool always = false)
{
	if (v.empty() && !always) return;
	putKey(out, field, 2);
	putVarint(out, v.size());
	out += v;
}

}

std::vector<std::string> ReadStreamFile(const std::string& filename)
{
	gzFile f = gzopen(filename.c_str(), "rb");
	if (!f) throw std::runtime_error("cannot open " + filename);
	std::string raw;
	char buf[1 << 16];
	int n;
	while ((n = gzread(f, buf, sizeof(buf))) > 0) raw.append(buf, (size_t)n);
	gzclose(f);
	std::vector<std::string> records;
	Reader r { (const uint8_t*)raw.data(), (const uint8_t*)raw.data() + raw.size() };
	while (!r.done())
	{
		uint64_t count = r.varint();
		for (uint64_t i = 0; i < count; i++) records.push_back(r.str());
	}
	return records;
}

void WriteStreamFile(const std::string& filename, const std::vector<std::string>& records)
{
	// one group holding every record, as write_buffered(..., 0) produces for the final GAM (Aligner.cpp:310-314);
	// an empty vector writes an empty gzip member
	std::string raw;
	if (!records.empty())
	{
		putVarint(raw, records.size());
		for (auto& rec : records)
		{
			putVarint(raw, rec.size());
			raw += rec;
		}
	}
	gzFile f = gzopen(filename.c_str(), "wb");
	if (!f) throw std::runtime_error("cannot open " + filename + " for writing");
	// gzwrite takes an unsigned length: pieces of at most 1 GiB, every result checked
	const size_t piece = (size_t)1 << 30;
	for (size_t off = 0; off < raw.size(); off += piece)
	{
		const unsigned len = (unsigned)std::min(piece, raw.size() - off);
		if (gzwrite(f, raw.data() + off, len) != (int)len)
		{
			gzclose(f);
			throw std::runtime_error("short write to " + filename);
		}
	}
	if (gzclose(f) != Z_OK) throw std::runtime_error("cannot finish writing " + filename);
}

void ReadGraphFile(const std::string& filename, std::vector<DirectedGraph::BiNode>& nodes, std::vector<DirectedGraph::BiEdge>& edges)
{
	for (auto& rec : ReadStreamFile(filename))
	{
		Reader g { (const uint8_t*)rec.data(), (const uint8_t*)rec.data() + rec.size() };
		while (!g.done())
		{
			uint64_t key = g.varint();
			int field = (int)(key >> 3), wt = (int)(key & 7);
			if (field == 1 && wt == 2)
			{
				Reader n = g.sub();
				DirectedGraph::BiNode node;
				node.id = 0;
				while (!n.done())
				{
					uint64_t k2 = n.varint();
					int f2 = (int)(k2 >> 3), w2 = (int)(k2 & 7);
					if (f2 == 1 && w2 == 2) node.sequence = n.str();
					else if (f2 == 2 && w2 == 2) node.name = n.str();
					else if (f2 == 3 && w2 == 0) node.id = (int64_t)n.varint();
					else n.skip(w2);
				}
				nodes.push_back(node);
			}
			else if (field == 2 && wt == 2)
			{
				Reader e = g.sub();
				DirectedGraph::BiEdge edge { 0, 0, false, false };
				while (!e.done())
				{
					uint64_t k2 = e.varint();
					int f2 = (int)(k2 >> 3), w2 = (int)(k2 & 7);
					if (f2 == 1 && w2 == 0) edge.from = (int64_t)e.varint();
					else if (f2 == 2 && w2 == 0) edge.to = (int64_t)e.varint();
					else if (f2 == 3 && w2 == 0) edge.from_start = e.varint() != 0;
					else if (f2 == 4 && w2 == 0) edge.to_end = e.varint() != 0;
					else e.skip(w2);
				}
				edges.push_back(edge);
			}
			else g.skip(wt);
		}
	}
}

std::string EncodeGraph(const std::vector<DirectedGraph::BiNode>& nodes, const std::vector<DirectedGraph::BiEdge>& edges)
{
	std::string out;
	for (auto& n : nodes)
	{
		std::string m;
		putBytes(m, 1, n.sequence);
		putBytes(m, 2, n.name);
		putInt(m, 3, n.id);
		putBytes(out, 1, m, true);
	}
	for (auto& e : edges)
	{
		std::string m;
		putInt(m, 1, e.from);
		putInt(m, 2, e.to);
		putBool(m, 3, e.from_start);
		putBool(m, 4, e.to_end);
		putBytes(out, 2, m, true);
	}
	return out;
}

static vg::Position decodePosition(Reader r)
{
	vg::Position p;
	while (!r.done())
	{
		uint64_t k = r.varint();
		int f = (int)(k >> 3), w = (int)(k & 7);
		if (f == 1 && w == 0) p.node_id = (int64_t)r.varint();
		else if (f == 2 && w == 0) p.offset = (int64_t)r.varint();
		else if (f == 4 && w == 0) p.is_reverse = r.varint() != 0;
		else r.skip(w);
	}
	return p;
}

static vg::Mapping decodeMapping(Reader r)
{
	vg::Mapping m;
	while (!r.done())
	{
		uint64_t k = r.varint();
		int f = (int)(k >> 3), w = (int)(k & 7);
		if (f == 1 && w == 2) m.position = decodePosition(r.sub());
		else if (f == 2 && w == 2)
		{
			Reader e = r.sub();
			vg::Edit edit;
			while (!e.done())
			{
				uint64_t k2 = e.varint();
				int f2 = (int)(k2 >> 3), w2 = (int)(k2 & 7);
				if (f2 == 1 && w2 == 0) edit.from_length = (int32_t)e.varint();
				else if (f2 == 2 && w2 == 0) edit.to_length = (int32_t)e.varint();
				else if (f2 == 3 && w2 == 2) edit.sequence = e.str();
				else e.skip(w2);
			}
			m.edit.push_back(edit);
		}
		else if (f == 5 && w == 0) m.rank = (int64_t)r.varint();
		else r.skip(w);
	}
	return m;
}

vg::Alignment DecodeAlignment(const std::string& msg)
{
	vg::Alignment a;
	Reader r { (const uint8_t*)msg.data(), (const uint8_t*)msg.data() + msg.size() };
	while (!r.done())
	{
		uint64_t k = r.varint();
		int f = (int)(k >> 3), w = (int)(k & 7);
		if (f == 1 && w == 2) a.sequence = r.str();
		else if (f == 2 && w == 2)
		{
			Reader p = r.sub();
			while (!p.done())
			{
				uint64_t k2 = p.varint();
				int f2 = (int)(k2 >> 3), w2 = (int)(k2 & 7);
				if (f2 == 2 && w2 == 2) a.path.mapping.push_back(decodeMapping(p.sub()));
				else p.skip(w2);
			}
		}
		else if (f == 3 && w == 2) a.name = r.str();
		else if (f == 6 && w == 0) a.score = (int32_t)r.varint();
		else if (f == 7 && w == 0) a.query_position = (int32_t)r.varint();
		else r.skip(w);
	}
	return a;
}

std::string EncodeAlignment(const vg::Alignment& aln)
{
	// fields in field-number order, as protobuf serialises them
	std::string out;
	putBytes(out, 1, aln.sequence);
	{
		std::string path;
		for (auto& m : aln.path.mapping)
		{
			std::string mm;
			{
				std::string pos;
				putInt(pos, 1, m.position.node_id);
				putInt(pos, 2, m.position.offset);
				putBool(pos, 4, m.position.is_reverse);
				putBytes(mm, 1, pos, true);   // the reference always allocates a Position (GraphAligner.h:804-806)
			}
			for (auto& e : m.edit)
			{
				std::string ee;
				putInt(ee, 1, e.from_length);
				putInt(ee, 2, e.to_length);
				putBytes(ee, 3, e.sequence);
				putBytes(mm, 2, ee, true);
			}
			putInt(mm, 5, m.rank);
			putBytes(path, 2, mm, true);
		}
		putBytes(out, 2, path, true);          // set_allocated_path: present even when empty (GraphAligner.h:788-789)
	}
	putBytes(out, 3, aln.name);
	putInt(out, 6, aln.score);
	putInt(out, 7, aln.query_position);
	return out;
}

std::vector<vg::Alignment> ReadAlignmentFile(const std::string& filename)
{
	std::vector<vg::Alignment> result;
	for (auto& rec : ReadStreamFile(filename)) result.push_back(DecodeAlignment(rec));
	return result;
}

void WriteAlignmentFile(const std::string& filename, const std::vector<vg::Alignment>& alns)
{
	std::vector<std::string> recs;
	for (auto& a : alns) recs.push_back(EncodeAlignment(a));
	WriteStreamFile(filename, recs);
}

}
