#include "ChunkReader.h"

#include <algorithm>
#include <cstring>

#include "GzipTextWriter.h"
#include "TextSource.h"

namespace seqpurge
{

// reads the files of one list, cuts the inflated text after every `pairs` records
void readerLoop(const std::vector<std::string>& files, int pairs, ChunkQueue& out, WorkerPool* pool)
{
	try
	{
		std::vector<uint8_t> buf((size_t)4 << 20);
		for (size_t fi = 0; fi < files.size(); ++fi)
		{
			std::unique_ptr<TextSource> src = openTextSource(files[fi], pool); // BGZF inputs are inflated by the pool, anything else by gzFile
			std::unique_ptr<TextChunk> cur(new TextChunk());
			cur->file_index = fi;
			long long lines = 0;   // complete lines in cur
			size_t line_len = 0;   // bytes of the current (incomplete) line
			const long long cut = 4ll * pairs;
			for (;;)
			{
				const size_t n = src->read(buf.data(), buf.size());
				if (n == 0) break;
				const uint8_t* p = buf.data();
				const uint8_t* end = p + n;
				const uint8_t* start = p; // first byte not yet appended to cur
				while (p < end)
				{
					const uint8_t* nl = (const uint8_t*)memchr(p, '\n', (size_t)(end - p));
					if (!nl)
					{
						line_len += (size_t)(end - p);
						break;
					}
					line_len += (size_t)(nl - p);
					if (lines & 1) cur->max_read_len = (int)std::min<size_t>(std::max<size_t>((size_t)cur->max_read_len, line_len), 1u << 30);
					line_len = 0;
					++lines;
					p = nl + 1;
					if (lines == cut)
					{
						cur->data.insert(cur->data.end(), start, p);
						cur->records = pairs;
						out.push(std::move(cur));
						cur.reset(new TextChunk());
						cur->file_index = fi;
						lines = 0;
						start = p;
					}
				}
				cur->data.insert(cur->data.end(), start, end);
			}
			src.reset();
			if (line_len > 0) // unterminated last line
			{
				if (lines & 1) cur->max_read_len = (int)std::min<size_t>(std::max<size_t>((size_t)cur->max_read_len, line_len), 1u << 30);
				++lines;
			}
			cur->records = (int)((lines + 3) / 4);
			cur->file_end = true;
			out.push(std::move(cur));
		}
		out.finish(nullptr);
	}
	catch (...)
	{
		out.finish(std::current_exception());
	}
}

// the `index`-th record of a chunk as the reference's reader delivers it (error reporting only)
FastqEntry entryAt(const TextChunk& c, int index)
{
	FastqEntry e;
	const uint8_t* p = c.data.data();
	const uint8_t* end = p + c.data.size();
	std::string* field[4] = {&e.header, &e.bases, &e.header2, &e.qualities};
	long long line = 0;
	while (p < end && line < 4ll * index + 4)
	{
		const uint8_t* nl = (const uint8_t*)memchr(p, '\n', (size_t)(end - p));
		const uint8_t* le = nl ? nl : end;
		if (line >= 4ll * index)
		{
			const uint8_t* e2 = le;
			while (e2 > p && e2[-1] == '\r') --e2;
			field[line - 4ll * index]->assign((const char*)p, (size_t)(e2 - p));
		}
		++line;
		p = nl ? nl + 1 : end;
	}
	return e;
}


} // namespace seqpurge
