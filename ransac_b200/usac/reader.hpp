// reader.hpp - the on-disk formats of the reference's datasets, read into the N x 4 (N x 2 for lines) float32 matrices the
// USAC classes take. Same static interface as the reference's Reader (detector/Reader.h:8-18); the line-fitting files are read
// the way ImageData does (dataset/GetImage.h:88-116). Bodies re-authored over plain streams (the shim cv::Mat has no push_back).
#pragma once
#include <fstream>
#include <sstream>
#include <string>
#include <vector>

#include "mat.hpp"

class Reader {
    static cv::Mat to_mat(const std::vector<float>& v, int cols) {
        cv::Mat m((int)(v.size() / cols), cols);
        for (size_t i = 0; i < v.size(); i++) m.ptr()[i] = v[i];
        return m;
    }
public:
    // "x1 y1 z1 x2 y2 z2 isinlier" per row (detector/Reader.cpp:13-48): two N x 2 matrices
    static void read_points(cv::Mat& pts1, cv::Mat& pts2, const std::string& filename) {
        std::fstream file(filename, std::ios_base::in);
        std::vector<float> a, b;
        float x1, y1, z1, x2, y2, z2, inl;
        while (file >> x1 >> y1 >> z1 >> x2 >> y2 >> z2 >> inl) { a.push_back(x1); a.push_back(y1); b.push_back(x2); b.push_back(y2); }
        pts1 = to_mat(a, 2); pts2 = to_mat(b, 2);
    }
    // the same rows: indices of the rows flagged as inliers (detector/Reader.cpp:76-118)
    static void getInliers(const std::string& filename, std::vector<int>& inliers) {
        std::fstream file(filename, std::ios_base::in);
        inliers.clear();
        float x1, y1, z1, x2, y2, z2;
        int inl, p = 0;
        while (file >> x1 >> y1 >> z1 >> x2 >> y2 >> z2 >> inl) { if (inl > 0) inliers.push_back(p); p++; }
    }
    // "x1 y1 1 x2 y2 1" per row -> N x 4 (detector/Reader.cpp:58-68)
    static void getPointsNby6(const std::string& filename, cv::Mat& points) {
        std::fstream file(filename, std::ios_base::in);
        std::vector<float> v;
        float x1, y1, z1, x2, y2, z2;
        while (file >> x1 >> y1 >> z1 >> x2 >> y2 >> z2) { v.push_back(x1); v.push_back(y1); v.push_back(x2); v.push_back(y2); }
        points = to_mat(v, 4);
    }
    // nine numbers, row-major (detector/Reader.cpp:129-146); the reference exits when the file is missing
    static void getMatrix3x3(const std::string& filename, cv::Mat& model) {
        model = cv::Mat(3, 3);
        std::fstream file(filename, std::ios_base::in);
        if (!file.is_open()) throw std::runtime_error("Wrong direction to matrix file! " + filename);
        for (int i = 0; i < 9; i++) file >> model.ptr()[i];
    }
    // twelve numbers, row-major 3 x 4 (detector/Reader.cpp:285-301)
    static void readProjectionMatrix(cv::Mat& P, const std::string& filename) {
        P = cv::Mat(3, 4);
        std::fstream file(filename, std::ios_base::in);
        if (!file.is_open()) throw std::runtime_error("Wrong direction to Projection matrix file! " + filename);
        for (int i = 0; i < 12; i++) file >> P.ptr()[i];
    }
    // "N" then N rows "x1 y1 x2 y2" (detector/Reader.cpp:182-210): the *_pts.txt files
    static bool LoadPointsFromFile(cv::Mat& points, const char* file) {
        std::ifstream infile(file);
        if (!infile.is_open()) return false;
        std::string line;
        int n = 0, row = -1;
        while (std::getline(infile, line)) {
            if (row < 0) { n = std::atoi(line.c_str()); if (n <= 0) return false; points = cv::Mat(n, 4); row = 0; continue; }
            if (row >= n) break;
            std::istringstream split(line);
            float* p = points.ptr() + (size_t)row * 4;
            split >> p[0] >> p[1] >> p[2] >> p[3];
            row++;
        }
        return row >= 0;
    }
    // the inverse (detector/Reader.cpp:150-180): all rows, or the rows listed in *inliers
    static bool SavePointsToFile(const cv::Mat& points, const char* file, std::vector<int>* inliers) {
        std::ofstream out(file, std::ios::out);
        if (!out.is_open()) return false;
        const int M = points.cols;
        const int count = inliers ? (int)inliers->size() : points.rows;
        out << count << std::endl;
        for (int i = 0; i < count; i++) {
            const float* p = points.ptr() + (size_t)(inliers ? inliers->at(i) : i) * M;
            for (int j = 0; j < M; j++) out << p[j] << " ";
            out << std::endl;
        }
        return true;
    }
    // EVD tentatives: a header line, then "x1,y1,x2,y2,FGINN_ratio,SNN_ratio,detector,descriptor,is_correct" (detector/Reader.cpp:216-269)
    static void readEVDPointsInliers(cv::Mat& points, std::vector<int>& inliers, const std::string& filename) {
        std::fstream file(filename, std::ios_base::in);
        inliers.clear();
        std::vector<float> v;
        std::string line, tok;
        std::getline(file, line);                                    // skip format
        int i = 0;
        while (std::getline(file, line)) {
            if (line.empty()) continue;
            std::istringstream row(line);
            std::string f[9];
            int k = 0;
            while (k < 9 && std::getline(row, tok, ',')) f[k++] = tok;
            if (k < 9) break;
            for (int c = 0; c < 4; c++) v.push_back(std::stof(f[c]));
            if (static_cast<bool>(std::stof(f[8]))) inliers.push_back(i);
            i++;
        }
        points = to_mat(v, 4);
    }
    // "N" then N indices (detector/Reader.cpp:303-320, the Strecha inlier lists)
    static void readInliers(std::vector<int>& inliers, const std::string& filename) {
        inliers.clear();
        std::fstream file(filename, std::ios_base::in);
        if (!file.is_open()) throw std::runtime_error("Wrong direction to inliers file! " + filename);
        int n = 0;
        file >> n;
        inliers.assign((size_t)std::max(n, 0), 0);
        for (int i = 0; i < n; i++) file >> inliers[i];
    }
    // line fitting sets: "width height noise a b c N" then N rows "x y" (dataset/GetImage.h:88-116); gt_model = (a, b, c)
    static bool readLine2d(cv::Mat& points, cv::Mat& gt_model, const std::string& filename) {
        std::ifstream file(filename);
        if (!file.is_open()) return false;
        float width, height, noise;
        int n = 0;
        gt_model = cv::Mat(1, 3);
        if (!(file >> width >> height >> noise >> gt_model.ptr()[0] >> gt_model.ptr()[1] >> gt_model.ptr()[2] >> n) || n <= 0) return false;
        points = cv::Mat(n, 2);
        for (int i = 0; i < 2 * n; i++) if (!(file >> points.ptr()[i])) return false;
        return true;
    }
};
