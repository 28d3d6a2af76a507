// EncodeRowsParity.cs -- checks K4os.Compression.LZ4's own C# engine against the known-answer rows
// this repository pins its GPU encoder to (tests/golden/encode_rows.json: length, Adler-32, SHA-256
// of LZ4Codec.Encode(L00_FAST) output, in the style of ChecksumBlockTests.cs:185-216).  The rows were
// produced by the upstream C engine the C# code is a port of; this test closes the loop on a machine
// that has .NET:   dotnet test bindings/csharp/ParityDump -e K4_ROWS=<repo>/tests/golden/encode_rows.json
//
// Inputs: only the generator kinds that need nothing but System.Random-free arithmetic are rebuilt
// here ("repeat", "lorem"); for the numpy-seeded kinds run `python tests/golden/dump_inputs.py <dir>`
// first and point K4_INPUTS at the directory (one file per row: <kind>-<size>-<seed>.bin).
// NOT COMPILED IN THIS REPOSITORY (no .NET SDK in the build image).
using System;
using System.IO;
using System.Linq;
using System.Security.Cryptography;
using System.Text.Json;
using K4os.Compression.LZ4;
using Xunit;

public class EncodeRowsParity
{
    private const string Lorem =
        "Sed ut perspiciatis unde omnis iste natus error sit voluptatem accusantium doloremque " +
        "laudantium, totam rem aperiam, eaque ipsa quae ab illo inventore veritatis et quasi " +
        "architecto beatae vitae dicta sunt explicabo. Nemo enim ipsam voluptatem quia voluptas " +
        "sit aspernatur aut odit aut fugit, sed quia consequuntur magni dolores eos qui ratione " +
        "voluptatem sequi nesciunt. Neque porro quisquam est, qui dolorem ipsum quia dolor sit amet. ";

    private static byte[] Input(string kind, int size, int seed)
    {
        var dir = Environment.GetEnvironmentVariable("K4_INPUTS");
        if (dir != null && File.Exists(Path.Combine(dir, $"{kind}-{size}-{seed}.bin")))
            return File.ReadAllBytes(Path.Combine(dir, $"{kind}-{size}-{seed}.bin"));
        if (kind == "repeat") return Enumerable.Repeat((byte)(seed & 0xFF), size).ToArray();
        if (kind == "lorem")
        {
            var one = System.Text.Encoding.ASCII.GetBytes(Lorem);
            var buf = new byte[size];
            for (var i = 0; i < size; i++) buf[i] = one[i % one.Length];
            return buf;
        }
        return null; // needs the dumped input file
    }

    private static uint Adler32(ReadOnlySpan<byte> data)
    {
        uint a = 1, b = 0;
        foreach (var x in data) { a = (a + x) % 65521; b = (b + a) % 65521; }
        return (b << 16) | a;
    }

    [Fact]
    public void CSharpEngineReproducesEveryPinnedRow()
    {
        var path = Environment.GetEnvironmentVariable("K4_ROWS") ?? "tests/golden/encode_rows.json";
        using var doc = JsonDocument.Parse(File.ReadAllText(path));
        var checkedRows = 0;
        foreach (var row in doc.RootElement.GetProperty("rows").EnumerateArray())
        {
            var kind = row.GetProperty("kind").GetString();
            var size = row.GetProperty("size").GetInt32();
            var seed = row.GetProperty("seed").GetInt32();
            var source = Input(kind, size, seed);
            if (source == null) continue;
            var target = new byte[LZ4Codec.MaximumOutputSize(size)];
            var n = LZ4Codec.Encode(source, target, LZ4Level.L00_FAST);
            Assert.Equal(row.GetProperty("len").GetInt32(), n);
            Assert.Equal(row.GetProperty("adler32").GetUInt32(), Adler32(target.AsSpan(0, n)));
            Assert.Equal(row.GetProperty("sha256").GetString(),
                Convert.ToHexString(SHA256.HashData(target.AsSpan(0, n))).ToLowerInvariant());
            foreach (var lim in row.GetProperty("limited").EnumerateArray())   // limitedOutput return codes
            {
                var cap = lim[0].GetInt32();
                var small = new byte[cap];
                Assert.Equal(lim[1].GetInt32(), LZ4Codec.Encode(source, small, LZ4Level.L00_FAST));
            }
            var back = new byte[size];
            Assert.Equal(size, LZ4Codec.Decode(target.AsSpan(0, n), back));
            Assert.True(back.AsSpan().SequenceEqual(source));
            checkedRows++;
        }
        Assert.True(checkedRows > 0);
    }
}
