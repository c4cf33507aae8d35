// swift-tools-version:6.0
// Drop-in replacement package for the decode hot path of tsolomko/SWCompression: same module name, same public types,
// bodies routed through the C ABI of libswcgpu.so (include/swcgpu.h).  COMPILE-UNVERIFIED: the build image has no Swift.
import PackageDescription

let package = Package(
    name: "SWCompression",
    products: [.library(name: "SWCompression", targets: ["SWCompression"])],
    targets: [
        .systemLibrary(name: "CSWCGPU", path: "Sources/CSWCGPU"),
        .target(name: "SWCompression", dependencies: ["CSWCGPU"], path: "Sources/SWCompression",
                linkerSettings: [.linkedLibrary("swcgpu")]),
    ],
    swiftLanguageModes: [.v6]
)
