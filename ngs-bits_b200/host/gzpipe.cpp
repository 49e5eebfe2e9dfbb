// gzpipe -- host-only test helper for the text streams of the command line: copies a (plain / gzip / BGZF) file through TextSource
// into a GzipTextWriter. Used by tests/test_text_streams.py (no GPU needed).
// usage: gzpipe <in> <out> [-threads N] [-bgzf] [-level L]
#include <cstdlib>
#include <cstring>
#include <iostream>
#include <memory>
#include <vector>

#include "GzipTextWriter.h"
#include "SeqPurgeTypes.h"
#include "TextSource.h"

using namespace seqpurge;

int main(int argc, char** argv)
{
	if (argc < 3)
	{
		std::cerr << "usage: gzpipe <in> <out> [-threads N] [-bgzf] [-level L]" << std::endl;
		return 2;
	}
	int threads = 1, level = 1;
	bool bgzf = false;
	for (int i = 3; i < argc; ++i)
	{
		if (!strcmp(argv[i], "-threads") && i + 1 < argc) threads = atoi(argv[++i]);
		else if (!strcmp(argv[i], "-level") && i + 1 < argc) level = atoi(argv[++i]);
		else if (!strcmp(argv[i], "-bgzf")) bgzf = true;
	}
	try
	{
		std::unique_ptr<WorkerPool> pool;
		if (threads > 1) pool.reset(new WorkerPool(threads));
		std::cerr << (pool && isBgzf(argv[1]) ? "parallel BGZF inflate" : "serial inflate") << std::endl;
		std::unique_ptr<TextSource> in = openTextSource(argv[1], pool.get());
		GzipTextWriter out(argv[2], level, pool.get(), bgzf);
		std::vector<uint8_t> buf((size_t)3 << 20);
		for (;;)
		{
			const size_t n = in->read(buf.data(), buf.size());
			if (n == 0) break;
			out.write(std::vector<uint8_t>(buf.begin(), buf.begin() + (long)n));
		}
		out.close();
	}
	catch (const Exception& e)
	{
		std::cerr << e.what() << std::endl;
		return 1;
	}
	return 0;
}
